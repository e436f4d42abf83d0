"""Parse the reference's regression tables (test/mpileup/*.reg) into runnable cases.

Each `P|F expected-file command` line (test/regression.sh:90-152) becomes a dict
{kind, expected, argv-ish shell command}.  `$samtools` is replaced by the tool
under test; `$fmt` lines are run for bam only (cram is out of scope).
"""
import re

def parse_reg(text):
    cases = []
    for ln in text.splitlines():
        if not ln or ln.startswith('#'):
            continue
        m = re.match(r'^(P|F|INIT)\s+(\S+)\s+(.*)$', ln)
        if not m:
            continue
        kind, exp, cmd = m.groups()
        if kind == 'INIT':
            continue
        cases.append(dict(kind=kind, expected=exp, cmd=cmd.strip()))
    return cases
