"""Pin the CPU oracle against every golden vector the reference holds for the
hot path: test/mpileup/{mpileup,depth}.reg, the test.pl
mpileup/coverage/large-position cases.  CPU only."""
import pytest
import golden_cases

CASES = golden_cases.all_cases()


@pytest.mark.parametrize('case', CASES, ids=[c['id'] for c in CASES])
def test_oracle_matches_reference_golden(case, oracle_bin, corpus):
    if case['skip']:
        pytest.skip(case['skip'])
    ok, out, err = golden_cases.run_case(case, oracle_bin, oracle_bin, corpus)
    assert ok, f"{case['cmd']}\nstderr: {err[-400:]!r}\nstdout head: {out[:300]!r}"
