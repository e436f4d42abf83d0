"""BAQ probe (development aid): stage an M Mb window with its FASTA (BAQ on every read), repeat the read stage on the
resident inputs, report the CUDA-event time of the BAQ kernels, reads/s, the non-fused FP64 rate on the algorithmic
operation count (l_qseq x 15 x 53) and a digest of the rewritten qualities (variants must agree).
  python tools/baq_probe.py [mb] [reps]     (B200_BAQ_REG=1 register kernel (default) | 0 warp kernel for every read)"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from samtools_b200 import engine, synth
mb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
soa = synth.make_region(int(mb * 1e6), seed=2, with_ref=True)
n = len(soa['pos'])
eng = engine.Engine(0)
eng.set_keep_raw(True)
conf = engine.default_stage_conf(engine.MODE_MPILEUP)
eng.stage(soa, conf)
ms = []
for _ in range(reps):
    eng.restage(); ms.append(eng.last_baq_ms)
q = eng.fetch_qual(int(soa['qual'].nbytes))
ops = float((soa['l_qseq'].astype(np.int64) * 15 * 53).sum())
m = float(np.median(ms))
print(f"BAQ_REG={os.environ.get('B200_BAQ_REG', '1')}  {mb:g} Mb  {n} reads  baq {m:8.3f} ms  {n / m / 1e3:8.2f} Mreads/s  "
      f"{ops / m / 1e9:7.3f} Top/s (algorithmic, non-FMA)  qual sha {hashlib.sha256(bytes(q)).hexdigest()[:16]}")
