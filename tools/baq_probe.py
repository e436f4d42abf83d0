"""BAQ-only probe (development aid): stage a 1 Mb window with BAQ twice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samtools_b200 import engine, synth
soa = synth.make_region(int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 1_000_000, seed=2, with_ref=True)
eng = engine.Engine(0)
for _ in range(2):
    eng.stage(soa, engine.default_stage_conf(engine.MODE_MPILEUP))
    print('stage ms', eng.last_stage_ms, 'reads', len(soa['pos']))
