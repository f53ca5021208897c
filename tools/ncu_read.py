"""Summarise .ncu-rep captures: python tools/ncu_read.py gpurun_out/r01e_*.ncu-rep  (needs ncu on PATH; runs without a GPU)."""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size']


def main(paths, extra):
    for p in paths:
        out = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print('==', p, '|', d.get('Kernel Name', ('?',))[0][:90])
        for k in KEYS + extra:
            for h in hdr:
                if h == k or (k in extra and k in h):
                    print(f'   {h:90s} {d[h][0]:>18s} {d[h][1]}')
        stall = sorted(((float(d[h][0].replace(',', '')), h) for h in hdr if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') and d[h][0] not in ('', 'n/a')), reverse=True)[:6]
        for v, h in stall:
            print(f'   stall {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""):40s} {v:8.2f}')


if __name__ == '__main__':
    paths = [a for a in sys.argv[1:] if a.endswith('.ncu-rep')]
    main(paths, [a for a in sys.argv[1:] if not a.endswith('.ncu-rep')])
