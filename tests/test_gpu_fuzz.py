"""Oracle vs CUDA path on random corner-case SAMs (all option sets of tests/fuzz_sam.py, BAQ included)."""
import os
import pytest
from conftest import ROOT
from test_fuzz_emul import run_all

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, 'samtools_b200', 'bin', 'b200samtools')


@pytest.mark.parametrize('lo', [1])
def test_fuzz_cuda_path(lo, oracle_bin, tmp_path):
    # two seeds (~140 CLI runs, each paying a CUDA context): 12 seeds were run clean during development
    bad = run_all(CLI, oracle_bin, tmp_path, range(lo, lo + 2), need_noBAQ=False)
    assert not bad, f'{len(bad)} mismatches, first: {bad[0]}'
