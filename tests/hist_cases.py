"""`coverage -m / -A / -D / -w` command lines shared by the emulation-harness and the GPU test.  The reference holds no
golden file for the histogram views (test/test.pl:4143-4160 covers the tabular mode only), so the product is compared
with the oracle's restatement of coverage.c:223-304, 609-660: parity unpinned."""
import os, subprocess

OPTS = ['-m -w 30', '-A -w 17', '-D -w 25', '-m -w 200', '-w 10 -r T1:5-30', '-m -w 12 -Q 8 -q 45 --min-depth 2', '-D -A -w 33 -r T2',
        '-m', '-D -o {out}']


def run_all(cli, oracle_bin, corpus, tmp_path, synth_sam=None, extra_env=None):
    """returns the list of (options, file) whose output differs"""
    sample = os.path.join(corpus, 'test', 'dat', 'sample.sam')
    files = [sample] + ([synth_sam] if synth_sam else [])
    env = dict(os.environ, COLUMNS='111')
    if extra_env:
        env.update(extra_env)
    bad = []
    for f in files:
        for k, o in enumerate(OPTS):
            if f != sample:                      # the synthetic contig is called chr1
                o = o.replace('T1:5-30', 'chr1:5001-20000').replace('-r T2', '-r chr1:100-1100')
            outs = []
            for tag, exe in (('p', cli), ('o', oracle_bin)):
                of = str(tmp_path / f'h{k}{tag}.txt')
                r = subprocess.run(f'{exe} coverage {o.format(out=of)} {f}', shell=True, capture_output=True, env=env)
                outs.append((r.returncode, r.stdout if '{out}' not in o else open(of, 'rb').read()))
            if outs[0] != outs[1] or outs[1][0] != 0 or len(outs[1][1]) < 100:
                bad.append((o, os.path.basename(f), outs[0][1][:300], outs[1][1][:300]))
    return bad
