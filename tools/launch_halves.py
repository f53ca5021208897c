"""Per-kernel totals of the SECOND half of an ncu launch list (a warm-up step + one step captured back to back):
    python tools/launch_halves.py gpurun_out/x_launches.csv [top]"""
import collections, csv, re, sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    ix = {h: i for i, h in enumerate(rows[hdr])}
    data = [(r[ix['Kernel Name']], float(r[ix['Metric Value']].replace(',', ''))) for r in rows[hdr + 1:]
            if len(r) > ix['Metric Value'] and r[ix['Metric Name']] == 'gpu__time_duration.sum']
    unit = rows[hdr + 1][ix['Metric Unit']]
    scale = {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0}.get(unit, 1e-6)
    half = data[len(data) // 2:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in half:
        n = re.sub(r'\(.*', '', n)[:72]
        agg[n][0] += 1
        agg[n][1] += t
    print(f'{path}: {len(half)} launches in the second half, {sum(v[1] for v in agg.values()) * scale:.3f} ms')
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{n:74s} {c:5d} {t * scale:8.3f}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
