"""File -> text timing of the CLI surface (third measurement level): the same synthetic
coordinate-sorted SAM through `b200samtools mpileup -a` (GPU engine behind the C ABI) and through the CPU oracle,
both writing to /dev/null.  This level is bound by host-side SAM/BAM decoding (single thread today),
not by the pileup path; it is a development aid, not the bench metric.

  python tools/cli_e2e.py [region_mb] [repeats]"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from samtools_b200 import synth


def main():
    mb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    cli = os.environ.get('B200_TEST_CLI') or os.path.join(ROOT, 'samtools_b200', 'bin', 'b200samtools')
    oracle = os.path.join(ROOT, 'oracle', '_build', 'plp_oracle')
    ncols = int(mb * 1e6)
    soa = synth.make_region(ncols, seed=2)
    with tempfile.TemporaryDirectory() as td:
        sam = os.path.join(td, 'r.sam')
        synth.write_sam(sam, soa)
        out = {'region_mb': mb, 'reads': len(soa['pos']), 'sam_bytes': os.path.getsize(sam)}
        for name, exe in (('b200samtools', cli), ('oracle', oracle)):
            best = None
            for _ in range(reps):
                t0 = time.perf_counter()
                with open(os.devnull, 'wb') as dn:
                    subprocess.run([exe, 'mpileup', '-a', sam], stdout=dn, check=True)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            out[name] = {'seconds': best, 'positions_per_s': ncols / best}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
