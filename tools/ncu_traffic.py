"""Record the DRAM traffic of the roofline kernels from an ncu --set full capture, keyed to the kernel sources it was taken
from, so that bench.py can report `roofline.traffic` only while those sources are unchanged (a stale number is worse than
null).   python tools/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/r02_traffic.json profiles/<summary kept next to it>"""
import csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ['samtools_b200/csrc/mpileup_ent.cuh', 'samtools_b200/csrc/mpileup_ss.cuh', 'samtools_b200/csrc/plp_core.h', 'samtools_b200/csrc/engine.cu']


def sources_digest():
    h = hashlib.sha256()
    for s in SOURCES:
        h.update(open(os.path.join(ROOT, s), 'rb').read())
    return h.hexdigest()


def main():
    rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ''
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('Kernel Name')
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    acc = {}
    for r in rows[2:]:
        name = r[ik].split('(')[0].replace('void ', '').split('<')[0]
        b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        acc.setdefault(name, []).append(b)
    rec = {'sources': SOURCES, 'sources_sha256': sources_digest(), 'capture': note,
           'dram_bytes_per_launch': {k: sum(v) / len(v) for k, v in acc.items()}}
    json.dump(rec, open(out, 'w'), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == '__main__':
    main()
