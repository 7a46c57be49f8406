"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/summarize_launches.py launches.csv [top] > profiles/rNN_launches.txt"""
import collections
import csv
import sys


def main():
  path = sys.argv[1]
  top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
  lines = [l for l in open(path, errors='replace') if l.startswith('"')]
  rows = list(csv.reader(lines))
  head = rows[0]
  iname, ival, iunit = head.index('Kernel Name'), head.index('Metric Value'), head.index('Metric Unit')
  imet = head.index('Metric Name')
  total, per = 0.0, collections.defaultdict(lambda: [0.0, 0])
  for r in rows[1:]:
    if r[imet] != 'gpu__time_duration.sum':
      continue
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}[r[iunit]]
    ms = float(r[ival].replace(',', '')) * scale
    per[r[iname]][0] += ms
    per[r[iname]][1] += 1
    total += ms
  print(f'# per-kernel totals over {sum(v[1] for v in per.values())} launches, top {top}; total {total:.1f} ms')
  for name, (ms, n) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100 * ms / total:5.1f}%  {ms:8.2f} ms  n={n:5d}  {name[:110]}')


if __name__ == '__main__':
  main()
