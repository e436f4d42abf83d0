"""The BAQ arithmetic the device kernel k_baq_reg executes (samtools_b200/csrc/baq_reg.h: band row in registers,
diagonal coordinates, in-place row updates) single-stepped on the CPU and compared, read by read, with the oracle's
restatement of sam_prob_realn + probaln_glocal -- bit-exact qualities required.  The same comparison runs on the GPU
through the golden BAQ cases (16/19/21/23/33/34.out) and the C2-size test of tests/test_gpu_parity.py."""
import os, subprocess
import pytest
from conftest import ROOT


@pytest.fixture(scope='module')
def baq_host():
    subprocess.run([os.path.join(ROOT, 'tests', 'emul', 'build.sh')], check=True)
    return os.path.join(ROOT, 'tests', 'emul', '_build', 'baq_host')


def test_register_band_baq_matches_oracle(baq_host, tmp_path):
    from samtools_b200 import synth
    for seed, maker in ((7, lambda: synth.make_batch(length=40_000, depth=30, seed=7)), (5, lambda: synth.make_region(60_000, seed=5, with_ref=True))):
        soa = maker()
        sam, fa = str(tmp_path / f's{seed}.sam'), str(tmp_path / f's{seed}.fa')
        synth.write_sam(sam, soa); synth.write_fasta(fa, soa['tid_name'], soa['ref_full'])
        r = subprocess.run([baq_host, sam, fa], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        n_fast = int(r.stdout.split('(')[1].split()[0])
        assert n_fast > 5000, r.stdout          # the register path must actually be exercised


def test_register_band_baq_on_fuzz_reads(baq_host, tmp_path):
    """corner-case CIGARs (clips, indels, pads, leading deletions), N bases, window clipped at both contig ends"""
    import fuzz_sam
    for seed in (1, 2, 3, 6, 7):
        sam, fa = fuzz_sam.make_sam(seed)
        (tmp_path / f'f{seed}.sam').write_text(sam); (tmp_path / f'f{seed}.fa').write_text(fa)
        r = subprocess.run([baq_host, str(tmp_path / f'f{seed}.sam'), str(tmp_path / f'f{seed}.fa')], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
