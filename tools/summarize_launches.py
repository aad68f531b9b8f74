"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import sys


def main(path, title):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg, n = collections.OrderedDict(), 0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = row['Kernel Name'].split('(')[0].replace('void ', '')
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print('# %s\n' % title)
    print('%d launches, %.2f ms summed kernel time (ncu: cold-cache, serialised -- compare SHARES, not absolutes)\n' % (n, tot / 1e3))
    print('| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.1f | %.2f | %.1f%% |' % (k, c, t, t / c, 100 * t / tot))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
