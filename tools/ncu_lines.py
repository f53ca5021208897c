"""Per-source-line hot spots of one .ncu-rep (captured with --import-source on, built with -lineinfo):
   python tools/ncu_lines.py gpurun_out/x.ncu-rep [top]"""
import csv, io, subprocess, sys


def main(path, top=25):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == 'Line No')
    i_s, i_e = hdr.index('# Samples'), hdr.index('Instructions Executed')
    lines, tot_s, tot_e, cur_file = [], 0, 0, ''
    for r in rows:
        if r and r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
        if len(r) > i_e and r[0] not in ('', 'Line No') and r[0].isdigit():
            try:
                s, e = int(r[i_s]), int(r[i_e])
            except ValueError:
                continue
            lines.append((s, e, cur_file, int(r[0]), r[1].strip()[:110]))
            tot_s += s; tot_e += e
    print(f'total samples {tot_s}, warp instructions {tot_e}')
    for s, e, f, n, src in sorted(lines, reverse=True)[:top]:
        print(f'{100 * s / max(tot_s, 1):5.1f}% smp {100 * e / max(tot_e, 1):5.1f}% inst  {f}:{n:<4d} {src}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
