"""Aggregate one training step of an ncu launch list (gpu__time_duration CSV) by kernel:
   python tools/launch_summary.py gpurun_out/x_launches.csv [other.csv]   (second file: side-by-side comparison)"""
import collections, csv, re, sys


def load(path):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith('==')))
    idx = [i for i, r in enumerate(rows) if 'sgd_kernel' in r['Kernel Name']]
    a, b = idx[1] + 1, idx[2] + 1           # the second full step in the capture
    agg = collections.OrderedDict()
    for r in rows[a:b]:
        n = re.sub(r'^void ', '', r['Kernel Name']).replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
        m = re.match(r'([\w:]+)(<[^(]*>)?', n)
        k = (m.group(1) + (m.group(2) or '')).replace('gemm::', '').replace('(int)', '').replace('(bool)', '')[:48]
        v = agg.setdefault(k, [0, 0.0])
        v[0] += 1; v[1] += float(r['Metric Value']) / 1e6
    return agg, b - a


def main(paths):
    aggs = [load(p) for p in paths]
    keys = sorted(set(k for a, _ in aggs for k in a), key=lambda k: -aggs[-1][0].get(k, [0, 0])[1])
    for (a, n), p in zip(aggs, paths):
        print(f'{p}: {n} launches, {sum(v[1] for v in a.values()):.3f} ms')
    for k in keys:
        print(f'{k:50s}' + ''.join(f' {a.get(k, [0, 0])[0]:4d} {a.get(k, [0, 0])[1]:8.3f}' for a, _ in aggs))


if __name__ == '__main__':
    main(sys.argv[1:])
