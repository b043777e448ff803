"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0; n = 0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    key = row['Kernel Name'].replace('void ', '').replace('unnamed>::', '').split('(')[0][:70] + ' grid=' + row['Grid Size']
    agg[key][0] += 1; agg[key][1] += v; tot += v; n += 1
print('launches', n, 'total us %.1f' % tot)
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%5.1f%% %8.1f us %5d  avg %7.1f  %s' % (100 * t / tot, t, c, t / c, k))
