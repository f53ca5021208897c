"""Condense the .ncu-rep files of tools/ncu_suite.sh into profiles/<tag>_ncu_full.json (read by bench.py for roofline.traffic):
    python tools/ncu_to_json.py r01t "<code state>"          (needs ncu on PATH; no GPU)"""
import csv, glob, io, json, subprocess, sys

B, C, T = 256, 96, 56 * 56
M = B * T
ALG = {'gemm_fc1': ("fc1 + GELU + GELU' (M=802816, N=384, K=96; bf16 in, two bf16 outputs)", M * C * 2 + 4 * C * C * 2 + 2 * M * 4 * C * 2),
       'gemm_qkv': ('qkv projection (M=802816, N=288, K=96)', M * C * 2 + 3 * C * C * 2 + M * 3 * C * 2),
       'gemm_outproj': ('attention out-projection + bias + residual (M=802816, N=96, K=96)', M * C * 2 * 3 + C * C * 2),
       'gemm_wgrad_small': ('out-projection weight gradient (96 x 96, contraction over 802816 tokens, split 147)', 2 * M * C * 2 + 147 * C * C * 4),
       'attn_bwd': ('window attention backward, stage 1 (49152 (window, head) tasks)', 9 * M * C * 2),
       'attn_fwd': ('window attention forward, stage 1', 4 * M * C * 2 + M * 3 * 4),
       'ln_fwd': ('LayerNorm forward C=96 (M=802816)', 2 * M * C * 2 + 2 * M * 4),
       'ln_bwd': ('LayerNorm backward C=96 (M=802816), residual input = dy in this driver', 3 * M * C * 2 + 2 * M * 4),
       'gallery': ('cosine filter 8192 queries x 262144 gallery rows, 512-d fp16', (8192 + 262144) * 512 * 2),
       'gallery_filter': ('cosine filter main pass, 50000 queries x 125000 gallery rows, 512-d fp16 (bench gallery leg)', (50000 + 125000) * 512 * 2),
       'gallery_rerank': ('exact fp64 re-rank of 128 candidates per query, 50000 queries (fp32 rows gathered)', 50000 * 129 * 512 * 4)}
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 's': 1}


def main(tag, state):
    out = {}
    for p in sorted(glob.glob(f'gpurun_out/{tag}_*.ncu-rep')):
        name = p.split(f'{tag}_')[1].replace('.ncu-rep', '')
        if name not in ALG:
            continue
        o = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(o)))
        d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
        val = lambda k: float(d[k].replace(',', '')) * SCALE.get(u[k], 1)
        rd, wr, t = val('dram__bytes_read.sum'), val('dram__bytes_write.sum'), val('gpu__time_duration.sum')
        out[name] = {'kernel': d['Kernel Name'][:100], 'what': ALG[name][0], 'duration_us': round(t * 1e6, 2), 'dram_bytes_read': rd,
                     'dram_bytes_write': wr, 'dram_bytes': rd + wr, 'algorithmic_bytes': ALG[name][1], 'dram_gbs': round((rd + wr) / t / 1e9, 1),
                     'issue_active_pct': round(val('smsp__issue_active.avg.pct_of_peak_sustained_active'), 1),
                     'tensor_pipe_pct': round(val('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active'), 1),
                     'registers': int(val('launch__registers_per_thread')), 'l2_hit_pct': round(val('lts__t_sector_hit_rate.pct'), 1),
                     'warp_instructions': val('sm__inst_executed.sum') if 'sm__inst_executed.sum' in d else None}
        print(name, out[name]['duration_us'], 'us', round(out[name]['dram_bytes'] / 1e6, 1), 'MB DRAM vs', round(ALG[name][1] / 1e6, 1), 'MB algorithmic')
    json.dump({'how': 'ncu --set full --clock-control none --import-source on, one launch per kernel at the bench stage-1 shapes (tools/ncu_suite.sh); '
                      'cold caches, serialised', 'code_state': state, 'kernels': out}, open(f'profiles/{tag}_ncu_full.json', 'w'), indent=1)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
