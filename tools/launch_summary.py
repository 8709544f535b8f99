"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) per kernel.
usage: python tools/launch_summary.py gpurun_out/launches.csv profiles/rNN_launches_summary_X.csv [steps]
`steps` = how many steps of the bench the capture window holds (the totals are divided by it)."""
import csv
import re
import sys


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'mv2d::', '', name)
    m = re.match(r'([\w:]+(<[^(]*>)?)', name)
    return m.group(1) if m else name[:60]


def main(src, dst, steps=1):
    rows = []
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            rows.append((short(r['Kernel Name']), float(r['Metric Value'].replace(',', '')) / 1e3))
    agg = {}
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches_per_step', 'total_us', 'share_pct', 'avg_us'])
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, round(n / steps, 2), round(us / steps, 1), round(100 * us / tot, 1), round(us / n, 1)])
        w.writerow(['TOTAL (ncu: serialised, cold cache)', round(len(rows) / steps, 1), round(tot / steps, 1), 100.0, ''])
    print(open(dst).read())


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
