"""Quick device-side timing of the column kernels (development aid).
  python tools/perf_probe.py [region_mb] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from samtools_b200 import engine, synth

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
soa = synth.make_region(int(mb * 1e6), seed=2)
ref = soa['ref_full']
ncols = int(mb * 1e6)
eng = engine.Engine(0)
bytes_in = synth.algorithmic_bytes_in(soa)


def run(tag, sconf, fn, with_ref):
    soa['ref'] = ref if with_ref else None
    st = eng.stage(soa, sconf)          # first call pays the buffer allocations
    st = eng.stage(soa, sconf)
    stage_ms = eng.last_stage_ms
    ms = []
    for _ in range(reps):
        n = fn()
        ms.append(eng.last_kernel_ms)
    k = float(np.median(ms))
    print(f'{tag:28s} stage {stage_ms:8.2f} ms   kernel {k:8.3f} ms   {ncols / k / 1e3:9.1f} Mcol/s   out {n / 1e6:8.1f} MB   '
          f'{(bytes_in + n) / k / 1e6:8.1f} GB/s', flush=True)


E = engine
run('mpileup -a (no ref)', E.default_stage_conf(E.MODE_MPILEUP), lambda: eng.mpileup_text(all=1, fetch=False), False)
run('mpileup -a -B -f', E.default_stage_conf(E.MODE_MPILEUP, baq=0), lambda: eng.mpileup_text(all=1, fetch=False), True)
run('mpileup -f (BAQ)', E.default_stage_conf(E.MODE_MPILEUP), lambda: eng.mpileup_text(fetch=False), True)
run('depth -a', E.default_stage_conf(E.MODE_DEPTH), lambda: eng.depth_text(all=1, fetch=False), False)
run('depth -a -q 20', E.default_stage_conf(E.MODE_DEPTH), lambda: eng.depth_text(all=1, min_qual=20, fetch=False), False)
st = eng.stage(soa, E.default_stage_conf(E.MODE_COVERAGE, end=ncols)); t = []
for _ in range(reps):
    eng.coverage(); t.append(eng.last_kernel_ms)
print(f'{"coverage":28s} stage {eng.last_stage_ms:8.2f} ms   kernel {np.median(t):8.3f} ms   {ncols / np.median(t) / 1e3:9.1f} Mcol/s')
