"""Golden-output cases taken from the reference's own test-suite.

The corpus lives in tests/golden/ref_corpus.tar.gz (built by
tests/golden/make_golden.py from /root/reference/test).  `prepare(tmp)` unpacks
it and replays the INIT lines of test/mpileup/{mpileup,depth}.reg that only
convert formats (the tools here sniff SAM vs BAM, so a copy is enough);
`all_cases()` lists every P/F line of the two regression tables plus the
test.pl mpileup / coverage / large-position cases.

A case is run with `run_case(case, tool, view_tool, root)`: `$samtools view`
inside a command is always served by the oracle's `view` (format plumbing, not
on the hot path); every other `$samtools` is the tool under test.
"""
import os, re, shutil, subprocess, tarfile, gzip

HERE = os.path.dirname(os.path.abspath(__file__))
BUNDLE = os.path.join(HERE, 'golden', 'ref_corpus.tar.gz')
from regtable import parse_reg


def prepare(tmp):
    root = os.path.join(str(tmp), 'corpus')
    if os.path.isdir(root):
        return root
    os.makedirs(root)
    with tarfile.open(BUNDLE) as tf:
        tf.extractall(root, filter='data')
    mp = os.path.join(root, 'test', 'mpileup')
    # INIT: expected/1.out from the packed columns 3-6 (mpileup.reg:27)
    with gzip.open(os.path.join(mp, 'expected', '1.out.f3-6.gz'), 'rt') as f, \
            open(os.path.join(mp, 'expected', '1.out'), 'w') as o:
        for i, ln in enumerate(f, 1):
            o.write(f'CHROMOSOME_I\t{i}\t{ln}')
    # INIT: samtools view -b x.sam > x.bam  (readers sniff the format: copy)
    for s in ['xx#depth1', 'xx#depth2', 'xx#depth3', 'overlap50', 'anomalous', 'indels']:
        shutil.copy(os.path.join(mp, s + '.sam'), os.path.join(mp, s + '.bam'))
    # test.pl test_mpileup: dat/mpileup.N.sam -> "bam" list
    dat = os.path.join(root, 'test', 'dat')
    with open(os.path.join(dat, 'mpileup.bam.list'), 'w') as o:
        for n in (1, 2, 3):
            o.write(os.path.join(dat, f'mpileup.{n}.sam') + '\n')
    # BASELINE config 1: examples/ex1.sam.gz is headerless; `samtools view -t ex1.fa.fai` is replaced by the
    # readers taking the contig list from <ref>.fai, which `samtools faidx` would have written
    ex = os.path.join(root, 'examples')
    rows, name, ln, lb, lw, start, off = [], None, 0, 0, 0, 0, 0
    for line in open(os.path.join(ex, 'ex1.fa'), 'rb'):
        if line.startswith(b'>'):
            if name:
                rows.append((name, ln, start, lb, lw))
            name, ln, lb, lw, start = line[1:].split()[0].decode(), 0, 0, 0, off + len(line)
        else:
            if lb == 0:
                lb, lw = len(line.rstrip(b'\n')), len(line)
            ln += len(line.rstrip(b'\n'))
        off += len(line)
    rows.append((name, ln, start, lb, lw))
    with open(os.path.join(ex, 'ex1.fa.fai'), 'w') as o:
        o.write(''.join('%s\t%d\t%d\t%d\t%d\n' % r for r in rows))
    return root


EX1_CMDS = ['mpileup -f ex1.fa ex1.sam.gz', 'mpileup -B -f ex1.fa ex1.sam.gz', 'mpileup -a -x -Q 20 -q 30 -f ex1.fa ex1.sam.gz',
            'mpileup -B -aa -s -O -f ex1.fa -r seq2:100-600 ex1.sam.gz']


def all_cases():
    cases = []
    with tarfile.open(BUNDLE) as tf:
        for reg in ('mpileup.reg', 'depth.reg'):
            text = tf.extractfile(f'test/mpileup/{reg}').read().decode()
            for i, c in enumerate(parse_reg(text)):
                cmd = c['cmd'].replace('$fmt', 'bam').replace('$awk', 'awk')
                skip = None
                if '--output-mods' in cmd:
                    skip = 'base modifications (-M) are out of scope (SURVEY 8f rank 3)'
                cid = f"{reg.split('.')[0]}:{i}:{c['expected']}"
                cases.append(dict(id=cid, cwd='test/mpileup', cmd=cmd, expected='test/mpileup/expected/' + c['expected'],
                                  kind=c['kind'], skip=skip, table=reg))
    # depth.reg INIT (line 64): header-only BAM
    for c in cases:
        if 'xx#depth-noreads.bam' in c['cmd']:
            c['cmd'] = '$samtools view -H -o xx#depth-noreads.bam xx#depth1.sam; ' + c['cmd']
    d = 'test/dat'
    cases += [
        dict(id='testpl:mpileup.out.1', cwd=d, kind='P', skip=None, table='test.pl', expected=f'{d}/mpileup.out.1',
             cmd='$samtools mpileup -b mpileup.bam.list -f mpileup.ref.fa -r17:100-150', stderr=f'{d}/mpileup.err.1'),
        dict(id='testpl:mpileup.out.3', cwd=d, kind='P', skip=None, table='test.pl', expected=f'{d}/mpileup.out.3',
             cmd='$samtools mpileup -B --ff 0x14 -f mpileup.ref.fa -r17:1050-1060 mpileup.1.sam | grep -v mpileup'),
        dict(id='testpl:mpileup.out.5', cwd=d, kind='P', skip=None, table='test.pl', expected=f'{d}/mpileup.out.5',
             cmd='$samtools mpileup ../mpileup/overlap.bam | grep 128814202'),
        dict(id='testpl:coverage.1', cwd=d, kind='P', skip=None, table='test.pl', expected='test/coverage/1.expected',
             cmd='$samtools coverage sample.sam'),
        dict(id='testpl:coverage.1b', cwd=d, kind='P', skip=None, table='test.pl', expected='test/coverage/1.expected',
             cmd='$samtools coverage --min-depth 1 sample.sam'),
        dict(id='testpl:coverage.2', cwd=d, kind='P', skip=None, table='test.pl', expected='test/coverage/2.expected',
             cmd='$samtools coverage --min-depth 2 sample.sam'),
        dict(id='testpl:coverage.3', cwd=d, kind='P', skip=None, table='test.pl', expected='test/coverage/3.expected',
             cmd='$samtools coverage --min-depth 2 -Q 8 -q 45 sample.sam'),
        dict(id='testpl:coverage.4', cwd=d, kind='P', skip=None, table='test.pl', expected='test/coverage/4.expected',
             cmd="sed '/A1/d' sample.sam > sample1.sam; $samtools coverage --min-depth 1 sample.sam sample1.sam"),
        dict(id='testpl:coverage.5', cwd=d, kind='P', skip=None, table='test.pl', expected='test/coverage/5.expected',
             cmd="sed '/A1/d' sample.sam > sample1.sam; $samtools coverage --min-depth 4 sample.sam sample1.sam"),
        # bedcov (test/test.pl:3817-3825): a further caller of the same pileup iterator (bedcov.c:303-331)
        dict(id='testpl:bedcov', cwd='test/bedcov', kind='P', skip=None, table='test.pl', expected='test/bedcov/bedcov.expected',
             cmd='$samtools bedcov bedcov.bed bedcov.bam'),
        dict(id='testpl:bedcov_j', cwd='test/bedcov', kind='P', skip=None, table='test.pl', expected='test/bedcov/bedcov_j.expected',
             cmd='$samtools bedcov -j bedcov.bed bedcov.bam'),
        dict(id='testpl:bedcov_gG', cwd='test/bedcov', kind='P', skip=None, table='test.pl', expected='test/bedcov/bedcov_gG.expected',
             cmd='$samtools bedcov -g512 -G2048 bedcov_gG.bed bedcov.bam'),
        dict(id='testpl:bedcov_c', cwd='test/bedcov', kind='P', skip=None, table='test.pl', expected='test/bedcov/bedcov_c.expected',
             cmd='$samtools bedcov -c bedcov_gG.bed bedcov.bam'),
        dict(id='testpl:large_pos.depth', cwd='test/large_pos', kind='P', skip=None, table='test.pl',
             expected='test/large_pos/depth.expected.out', cmd='$samtools depth longref.sam'),
        dict(id='testpl:large_pos.depth_bed', cwd='test/large_pos', kind='P', skip=None, table='test.pl',
             expected='test/large_pos/depth_bed.expected.out', cmd='$samtools depth -b test.bed longref.sam'),
    ]
    return cases


def run_case(case, tool, view_tool, root, timeout=300):
    """Returns (ok, stdout, stderr).  ok already accounts for F (expected-fail) lines."""
    cmd = re.sub(r'\$samtools\s+view', view_tool + ' view', case['cmd'])
    cmd = cmd.replace('$samtools', tool)
    r = subprocess.run(cmd, shell=True, cwd=os.path.join(root, case['cwd']), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=timeout)
    exp = open(os.path.join(root, case['expected']), 'rb').read()
    out = r.stdout.replace(b'\r', b'')
    ok = out == exp
    if ok and case.get('stderr'):
        ok = r.stderr == open(os.path.join(root, case['stderr']), 'rb').read()
    if case['kind'] == 'F':
        ok = not ok
    return ok, out, r.stderr


# htslib-compatible iterator tier (T1): plp_dump client vs the oracle's `pileup-dump` (same lists for the CUDA engine and the
# emulation harness)
COMPAT_CASES = [
    ([], ['test/mpileup/mp_DI.sam']), ([], ['test/mpileup/mp_P.sam']), ([], ['test/mpileup/mp_N2.sam']),
    (['-o'], ['test/mpileup/overlap50.sam']), ([], ['test/mpileup/mpileup.1.bam']), (['-o'], ['test/mpileup/mpileup.1.bam']),
    ([], ['test/mpileup/xx#depth1.sam', 'test/mpileup/xx#depth2.sam']), (['-d', '8500'], ['test/mpileup/deep.sam']),
    ([], ['test/dat/mpileup.1.sam', 'test/dat/mpileup.2.sam', 'test/dat/mpileup.3.sam']),
]
COMPAT_IDS = ['DI', 'P', 'N2', 'overlap50', 'mpileup1', 'mpileup1-overlaps', 'two-files', 'maxcnt8500', 'three-files']
