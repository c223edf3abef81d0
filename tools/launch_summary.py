"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel/grid totals + one iteration."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    out = []
    for row in rows:
        t = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        t = t / 1000 if u == 'ns' else (t * 1000 if u == 'ms' else t)
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('mftb::', '')[:40]
        out.append((name, row['Grid Size'], t))
    return out


if __name__ == '__main__':
    rows = load(sys.argv[1])
    agg = collections.OrderedDict()
    tot = 0
    for name, grid, t in rows:
        a = agg.setdefault((name, grid), [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    print(f'{len(rows)} launches, total {tot:.1f} us')
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 24]:
        print(f'{t:9.1f} us {100 * t / tot:5.1f}%  n={n:3d}  avg={t / n:7.1f}  {k[0]} {k[1]}')
    if len(sys.argv) > 3:
        a, b = int(sys.argv[3]), int(sys.argv[4])
        for i in range(a, b):
            print(i, f'{rows[i][2]:7.1f}', rows[i][0], rows[i][1])
