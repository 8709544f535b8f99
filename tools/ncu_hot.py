"""Stall-sample histogram of one kernel from an .ncu-rep source page (first launch only)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
win = int(sys.argv[3]) if len(sys.argv) > 3 else 60
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
blocks = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
rows = rows[blocks[0]:(blocks[1] if len(blocks) > 1 else len(rows))]
hdr = rows[1]
si, src, ie = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Source'), hdr.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [(int(r[si] or 0), r[src].strip(), int(r[ie] or 0), r) for r in rows[2:] if len(r) > si]
tot = sum(d[0] for d in data)
print(kern, 'samples', tot, 'SASS instrs', len(data), 'warp-instrs executed', sum(d[2] for d in data))
agg = {}
for d in data:
    for i, h in stall_cols:
        agg[h] = agg.get(h, 0) + int(d[3][i] or 0)
print('  stall reasons:', ', '.join(f'{h[6:]} {100 * v / max(tot, 1):.0f}%' for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
for a in range(0, len(data), win):
    s = sum(d[0] for d in data[a:a + win]); n = sum(d[2] for d in data[a:a + win])
    if s * 100 > tot:
        print(f'  [{a:4d}] {100 * s / tot:5.1f}%  executed {n:8d}  {data[a][1][:60]}')
for i, d in sorted(enumerate(data), key=lambda x: -x[1][0])[:10]:
    print(f'   top [{i}] {100 * d[0] / tot:.1f}% x{d[2]} {d[1][:80]}')
