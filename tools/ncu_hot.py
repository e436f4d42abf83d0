"""Development aid: per-region instruction / stall-sample shares of one kernel from an .ncu-rep with source info.
  python tools/ncu_hot.py rep.ncu-rep kernel_regex [region_size]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
step = int(sys.argv[3]) if len(sys.argv) > 3 else 50
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if 'Address' in r and 'Source' in r)
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith('0x')]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iex]) for r in data); ts = sum(int(r[ismp]) for r in data)
print('kernel', kern, 'warp instructions', tot, 'samples', ts, 'SASS lines', len(data))
base = int(data[0][ia], 16)
acc = sacc = start = 0
for k, r in enumerate(data):
    acc += int(r[iex]); sacc += int(r[ismp])
    if (k + 1) % step == 0 or k == len(data) - 1:
        if acc > tot * 0.01 or sacc > ts * 0.01:
            print(f'{start:5d}-{k:5d} off {int(data[start][ia], 16) - base:#7x}: instr {acc / tot * 100:5.1f}%  samples {sacc / ts * 100:5.1f}%   {data[start][isrc].strip()[:70]}')
        acc = sacc = 0; start = k + 1
print('--- top stall instructions')
for r in sorted(data, key=lambda r: -int(r[ismp]))[:20]:
    print(f"{int(r[ia], 16) - base:#7x} smp {int(r[ismp]) / ts * 100:4.1f}% ex {int(r[iex]) / tot * 100:4.2f}%  {r[isrc].strip()[:90]}")
