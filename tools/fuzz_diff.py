"""Development aid: run the GPU fuzz cases of tests/test_fuzz_emul.run_all for some seeds / one command and print,
for every mismatch, the first differing output line of the oracle and of the CLI.
  python tools/fuzz_diff.py [--cmd gl] [--seeds 1,2] """
import argparse, os, subprocess, sys, tempfile, pathlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import fuzz_sam  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--cmd', default='gl')
ap.add_argument('--seeds', default='1,2,3,4')
a = ap.parse_args()
oracle = os.path.join(ROOT, 'oracle', '_build', 'plp_oracle')
cli = os.path.join(ROOT, 'samtools_b200', 'bin', 'b200samtools')
opts = {'gl': fuzz_sam.GL_OPTS, 'mpileup': fuzz_sam.MPILEUP_OPTS, 'depth': fuzz_sam.DEPTH_OPTS, 'coverage': fuzz_sam.COVERAGE_OPTS}[a.cmd]
shown = 0
for seed in [int(s) for s in a.seeds.split(',')]:
    with tempfile.TemporaryDirectory() as td:
        d = pathlib.Path(td)
        sam, fa = fuzz_sam.make_sam(seed)
        (d / 'x.sam').write_text(sam); (d / 'x.fa').write_text(fa)
        (d / 'x.bed').write_text('c0\t40\t300\nc0\t250\t600\nc1\t100\nc1\t95\t140\n'); (d / 'rg.txt').write_text('g2\n')
        (d / 'x2.sam').write_text(fuzz_sam.make_sam(seed + 100000, n_reads=25)[0])
        for opt in opts:
            opt = opt.format(bed='x.bed', rg='rg.txt')
            files = 'x.sam x2.sam' if (seed % 3 == 0 and '-r' not in opt) else 'x.sam'
            ref = '-f x.fa' if a.cmd in ('mpileup', 'gl') and seed % 4 != 1 else ''
            line = f'{a.cmd} {opt} {ref} {files}'
            A = subprocess.run(f'{oracle} {line}', shell=True, cwd=d, capture_output=True)
            B = subprocess.run(f'{cli} {line}', shell=True, cwd=d, capture_output=True)
            if A.stdout == B.stdout and (A.returncode == 0) == (B.returncode == 0):
                continue
            la, lb = A.stdout.split(b'\n'), B.stdout.split(b'\n')
            print(f'seed {seed}: {line}: rc {A.returncode}/{B.returncode}, {len(la)}/{len(lb)} lines; stderr {B.stderr[-200:]!r}')
            for i in range(max(len(la), len(lb))):
                x = la[i] if i < len(la) else b'<none>'; y = lb[i] if i < len(lb) else b'<none>'
                if x != y:
                    print('  line', i, '\n   oracle:', x.decode()[:600], '\n   cuda:  ', y.decode()[:600])
                    # the reads over that column
                    try:
                        pos = int(x.split(b'\t')[1]); name = x.split(b'\t')[0].decode()
                        for r in sam.split('\n'):
                            f = r.split('\t')
                            if len(f) > 10 and f[2] == name and int(f[3]) <= pos < int(f[3]) + 400:
                                print('     read:', '\t'.join(f[:9]), f[9][:20], f[10][:20])
                    except Exception:
                        pass
                    break
            shown += 1
            if shown >= 6:
                sys.exit(0)
