// Fused (shifted-)window attention, forward and backward, for the berniwal-variant Swin block
// (reference models/swin.py:101-135) on the 5th-generation tensor cores: every product is a tcgen05.mma with TMEM
// accumulators, q / k / v tiles arrive by TMA.
//
// Operand layout.  The LayerNorm in front of to_qkv writes its rows in WINDOW-MAJOR order (common.cuh: WinMap - the cyclic
// shift by 3 of the shifted blocks and the window partition are folded into that row permutation), so the output of the qkv
// GEMM, [tokens, 3C] = [q | k | v] with (head, dim) inside each part, holds every window as 49 consecutive rows.  One TMA
// box = [49 token rows x 64 columns] = one window x two heads of q, k or v, landing as a 128-B-row SWIZZLE_128B tile - the
// layout the UMMA shared-memory descriptors of the GEMM kernels already use.
//
// Work unit = (pair of consecutive windows, pair of heads).  A tile holds window A in rows 0..48 and window B in rows
// 64..112 (rows 49..63 / 113..127 stay zero), so all 128 TMEM lanes of an M = 128 MMA are used: lane r < 64 is token r of
// window A, lane 64 + r token r of window B.  Per (window pair, head):
//
//   forward   S  = Q K^T            one MMA series, N = 128 keys (A's then B's; a row uses the 64 columns of its own window)
//             P  = softmax(scale S + relpos + shift masks)   one thread per row: no shuffles; P -> smem (bf16, K-major)
//             O  = P V              two series (V of window A / of window B as the MN-major B operand, N = 32)
//   backward  S, dP = dO V^T        two series, N = 128
//             P  = exp(scale S + bias - lse),  D = rowsum(P o dP),  dS = P o (dP - D)        (P, dS -> smem)
//             dQ = dS K             (A = dS K-major,  B = K MN-major)        dK = dS^T Q,  dV = P^T dO
//             (A = the dS / P tile read MN-major: its two 64-row halves are exactly the two M chunks of an M = 128 operand)
//
// Warp roles: warp 0 = TMA producer, warp 1 = the single MMA-issuing thread, (backward: warps 2-3 gather the dO rows, which
// arrive in raster order, with cp.async), then two groups of four warps (one warp per TMEM lane quarter).  The heads of the
// stream of work units are dealt alternately to the two groups, each with its own TMEM accumulators and P / dS tiles, so one
// group's softmax overlaps the other's MMAs and global stores.
//
// The forward output and the backward's dO are in RASTER order (they meet the residual stream through to_out); dqkv and the
// row log-sum-exp are window-major like qkv.
#include "common.cuh"

#include "b200_fe.h"
#include "gemm_core.cuh"

namespace {

constexpr int kWs = 7;
constexpr int kWt = 49;
constexpr int kHd = 32;
constexpr int kBins = 169;
constexpr int kTile = 128 * 128;                   // bytes of one operand tile: [2 windows x 64 rows] x 128 B
constexpr int kWinOff = 64 * 128;                  // window B starts at row 64
constexpr int kBoxBytes = kWt * 128;               // bytes one TMA box delivers
constexpr int kBiasPitch = 52;                     // floats per row of the 49 x 49 bias table (conflict-free float4 rows)
constexpr int kBiasBytes = 10240;                  // 49 * 52 * 4 = 10192, rounded
constexpr int kBarBytes = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
// bit j set <=> window column (j % 7) >= 4 / window row (j / 7) >= 4: token j lies in the wrapped part under the mask
constexpr unsigned long long kColHi = 0x1C3870E1C3870ULL;
constexpr unsigned long long kRowHi = 0x1FFFFF0000000ULL;
constexpr unsigned long long kAll49 = 0x1FFFFFFFFFFFFULL;

struct AttnArgs {
  bf16* out;            // fwd: [B*H*W, C] attention output, raster order (input of to_out)
  float* lse;           // [B*H*W (window-major), heads] row log-sum-exp (fwd: written, nullable; bwd: read)
  const float* pos;     // [13*13] relative position table of this block
  const bf16* dout;     // bwd: grad wrt the attention output [B*H*W, C], raster order
  bf16* dqkv;           // bwd: [B*H*W (window-major), 3C]
  float* dpos_partial;  // bwd: [gridDim.x, 169]
  int B, H, W, C, heads, shifted;
  int nwin, nhp, units;       // windows, head pairs per window, work units = ceil(nwin / 2) * nhp
  float scale;
  FastDiv div_nhp, div_nww, div_nwh;
  uint32_t idesc_s, idesc_kmn, idesc_mnmn;    // N = 128 K-major x K-major; N = 32 K-major x MN-major; N = 32 MN-major x MN-major
  uint32_t wait_ns;
};

inline uint32_t attn_idesc(int n, bool a_mn, bool b_mn) {      // bf16 operands, fp32 accumulate, M = 128
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (a_mn ? 1u : 0u) << 15;
  d |= (b_mn ? 1u : 0u) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(128 >> 4) << 24;
  return d;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float fast_exp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }
__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 49 x 49 additive score table shared by every window and head of the block (models/swin.py:117-118), in the log2 domain:
// bias[i][j] = log2(e) * pos[r_j - r_i + 6][c_j - c_i + 6]
__device__ __forceinline__ void build_bias_table(float* bias_s, const float* __restrict__ pos) {
  for (int idx = threadIdx.x; idx < kWt * kBiasPitch; idx += blockDim.x) {
    const int i = idx / kBiasPitch, j = idx - i * kBiasPitch;
    float v = 0.f;
    if (j < kWt) {
      const int ri = i / kWs, ci = i - ri * kWs, rj = j / kWs, cj = j - rj * kWs;
      v = kLog2e * __ldg(pos + (rj - ri + kWs - 1) * (2 * kWs - 1) + (cj - ci + kWs - 1));
    }
    bias_s[idx] = v;
  }
}

// what one thread of a softmax group knows about "its" token: TMEM lane / tile row `row` = 64 * window slot + token
struct RowCtx {
  int row, ws, i;                 // tile row, window slot (0 = window A, 1 = B), token index (valid < 49)
  bool row_hi, col_hi;            // token in the wrapped rows / columns of a shifted window
  // per work unit
  bool valid;                     // a real token of a real window
  long long wm_row;               // window-major token row
  long long gr;                   // raster token row
  unsigned long long mask;        // shift-mask bits over the 49 keys (models/swin.py:49-62, :122-124), 0 when not flagged
};

__device__ __forceinline__ void row_unit(RowCtx& rc, const AttnArgs& a, int pair) {
  const int win = 2 * pair + rc.ws;
  rc.valid = win < a.nwin && rc.i < kWt;
  const int nww = a.W / kWs, nwh = a.H / kWs;
  const uint32_t w = static_cast<uint32_t>(min(win, a.nwin - 1));
  const uint32_t wrow = a.div_nww.div(w);
  const int wx = static_cast<int>(w - wrow * static_cast<uint32_t>(nww));
  const uint32_t img = a.div_nwh.div(wrow);
  const int wy = static_cast<int>(wrow - img * static_cast<uint32_t>(nwh));
  const int ic = min(rc.i, kWt - 1);
  const int r = ic / kWs, c = ic - r * kWs;
  const int off = a.shifted ? kWs / 2 : 0;
  int y = wy * kWs + r + off, x = wx * kWs + c + off;
  if (y >= a.H) y -= a.H;
  if (x >= a.W) x -= a.W;
  rc.gr = (static_cast<long long>(img) * a.H + y) * a.W + x;
  rc.wm_row = static_cast<long long>(w) * kWt + ic;
  unsigned long long m = 0ULL;
  if (a.shifted) {
    if (wy == nwh - 1) m |= rc.row_hi ? (~kRowHi & kAll49) : kRowHi;
    if (wx == nww - 1) m |= rc.col_hi ? (~kColHi & kAll49) : kColHi;
  }
  rc.mask = m;
}

// 16 consecutive accumulator columns -> scores in the log2 domain (scale * acc + bias), masked
template <int J0, int N>
__device__ __forceinline__ void scores_chunk(const uint32_t* r, const float* bias_row, float sc2, float sub, unsigned long long mask,
                                             float* s) {
#pragma unroll
  for (int j4 = 0; j4 < (N + 3) / 4; ++j4) {
    const float4 b = *reinterpret_cast<const float4*>(bias_row + J0 + 4 * j4);
    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (4 * j4 + e < N) s[4 * j4 + e] = fmaf(__uint_as_float(r[4 * j4 + e]), sc2, bb[e] - sub);
  }
  if (mask != 0ULL) {                    // warp-uniform: a warp's 32 rows belong to one window
#pragma unroll
    for (int j = 0; j < N; ++j)
      if ((mask >> (J0 + j)) & 1ULL) s[j] = -INFINITY;
  }
}

// one row of a P / dS tile: 49 values -> bf16, chunks 0..6 of the 128-B swizzled row (chunk 7 and keys 49..55 stay zero)
__device__ __forceinline__ void store_row_bf16(uint32_t tile_row_addr, int row, const float* v) {
#pragma unroll
  for (int c = 0; c < 6; ++c)
    st_shared_v4(tile_row_addr + ((c ^ (row & 7)) << 4), pack_bf16(v[8 * c], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]),
                 pack_bf16(v[8 * c + 4], v[8 * c + 5]), pack_bf16(v[8 * c + 6], v[8 * c + 7]));
  st_shared_v4(tile_row_addr + ((6 ^ (row & 7)) << 4), pack_bf16(v[48], 0.f), 0u, 0u, 0u);
}

// 32 fp32 accumulator columns of this lane -> bf16 * f -> 64 B of global memory
__device__ __forceinline__ void store_row32(bf16* dst, uint32_t taddr, float f, bool live) {
  uint32_t r0[16], r1[16];
  tmem_ld16(taddr, r0);
  tmem_ld16(taddr + 16, r1);
  tmem_ld_wait();
  if (live) {
    st_global_v4(dst, pack_bf16(__uint_as_float(r0[0]) * f, __uint_as_float(r0[1]) * f), pack_bf16(__uint_as_float(r0[2]) * f, __uint_as_float(r0[3]) * f),
                 pack_bf16(__uint_as_float(r0[4]) * f, __uint_as_float(r0[5]) * f), pack_bf16(__uint_as_float(r0[6]) * f, __uint_as_float(r0[7]) * f));
    st_global_v4(dst + 8, pack_bf16(__uint_as_float(r0[8]) * f, __uint_as_float(r0[9]) * f), pack_bf16(__uint_as_float(r0[10]) * f, __uint_as_float(r0[11]) * f),
                 pack_bf16(__uint_as_float(r0[12]) * f, __uint_as_float(r0[13]) * f), pack_bf16(__uint_as_float(r0[14]) * f, __uint_as_float(r0[15]) * f));
    st_global_v4(dst + 16, pack_bf16(__uint_as_float(r1[0]) * f, __uint_as_float(r1[1]) * f), pack_bf16(__uint_as_float(r1[2]) * f, __uint_as_float(r1[3]) * f),
                 pack_bf16(__uint_as_float(r1[4]) * f, __uint_as_float(r1[5]) * f), pack_bf16(__uint_as_float(r1[6]) * f, __uint_as_float(r1[7]) * f));
    st_global_v4(dst + 24, pack_bf16(__uint_as_float(r1[8]) * f, __uint_as_float(r1[9]) * f), pack_bf16(__uint_as_float(r1[10]) * f, __uint_as_float(r1[11]) * f),
                 pack_bf16(__uint_as_float(r1[12]) * f, __uint_as_float(r1[13]) * f), pack_bf16(__uint_as_float(r1[14]) * f, __uint_as_float(r1[15]) * f));
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr int kFwdStages = 3;
constexpr int kFwdStageBytes = 3 * kTile;                                  // q, k, v tiles
constexpr int kFwdThreads = 64 + 2 * 128;
constexpr int kFwdTileBytes = kFwdStages * kFwdStageBytes + 2 * kTile;      // operand ring + one P tile per group
constexpr int kFwdSmem = kFwdTileBytes + kBiasBytes + kBarBytes + 1024;
// TMEM columns of group g: S at 256 g (128 columns), O of window A / B at 256 g + 128 / + 160
constexpr int kFwdTmemO = 128;

__global__ void __launch_bounds__(kFwdThreads, 1) window_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t p_base = base + kFwdStages * kFwdStageBytes;
  float* bias_s = reinterpret_cast<float*>(base_ptr + kFwdTileBytes);
  const uint32_t bar_base = base + kFwdTileBytes + kBiasBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kFwdStages + s); };
  auto sfull_bar = [&](int g) { return bar_base + 8u * (2 * kFwdStages + g); };
  auto pfull_bar = [&](int g) { return bar_base + 8u * (2 * kFwdStages + 2 + g); };
  auto ofull_bar = [&](int g) { return bar_base + 8u * (2 * kFwdStages + 4 + g); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kFwdStages + 6);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kFwdTileBytes + kBiasBytes + 8 * (2 * kFwdStages + 6));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows 49..63 / 113..127 of every tile (and key columns 49..63 of the P tiles) are never written again: they must be zero
  for (int i = threadIdx.x; i < kFwdTileBytes / 16; i += kFwdThreads) st_shared_v4(base + 16u * i, 0u, 0u, 0u, 0u);
  build_bias_table(bias_s, a.pos);      // pos is a parameter: not produced by the preceding kernel
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int s = 0; s < kFwdStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(sfull_bar(g), 1); mbar_init(pfull_bar(g), 1); mbar_init(ofull_bar(g), 1); }
    mbar_fence_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();             // the zero fill is visible to TMA writes and UMMA reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_grid_sync();

  const int nhp = a.nhp;
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
        const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
        const int hp = unit - pair * nhp;
        mbar_wait_backoff(empty_bar(stage), phase ^ 1u, a.wait_ns);
        mbar_arrive_expect_tx(full_bar(stage), 6u * kBoxBytes);
        const uint32_t sb = base + stage * kFwdStageBytes;
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int w = 0; w < 2; ++w)
            tma_load_2d(sb + t * kTile + w * kWinOff, &tmap_qkv, full_bar(stage), t * a.C + hp * 64, (2 * pair + w) * kWt);
        if (++stage == kFwdStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      struct Pend { int valid, stage, slot, last; } pend[2] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
      uint32_t np[2] = {0u, 0u};                 // P tiles consumed per group (parity of p_full)
      auto issue_pv = [&](int g) {
        const Pend pd = pend[g];
        mbar_wait_backoff(pfull_bar(g), np[g] & 1u, a.wait_ns);
        ++np[g];
        tc_fence_after();
        const uint32_t vs = base + pd.stage * kFwdStageBytes + 2 * kTile;
        const uint64_t da = make_sw128_desc(p_base + g * kTile, 16, 1024);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          // B = V of window w, MN-major: contraction over its 64 key rows (8-row groups 1024 B apart, 16 rows per MMA),
          // N = the head's 32 columns inside the 64-column row
          const uint64_t db = make_sw128_desc(vs + w * kWinOff + 64 * pd.slot, 8192, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + g * 256 + kFwdTmemO + 32 * w, da + 2u * k, db + 128u * k, a.idesc_kmn, k > 0 ? 1u : 0u);
        }
        umma_commit(ofull_bar(g));
        if (pd.last) umma_commit(empty_bar(pd.stage));       // every MMA that reads this unit's tiles has been issued
        pend[g].valid = 0;
      };
      int k = 0;
      for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
        const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
        const int hp = unit - pair * nhp;
        const int nslots = min(2, a.heads - 2 * hp);
        for (int slot = 0; slot < nslots; ++slot, ++k) {
          const int g = k & 1;
          if (pend[g].valid) issue_pv(g);        // O of this group's previous head first: it frees S, and (last slot) the stage
          if (slot == 0) {
            mbar_wait_backoff(full_bar(stage), phase, a.wait_ns);
            tc_fence_after();
          }
          const uint32_t qs = base + stage * kFwdStageBytes, ks = qs + kTile;
          const uint64_t da = make_sw128_desc(qs + 64 * slot, 16, 1024);
          const uint64_t db = make_sw128_desc(ks + 64 * slot, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) umma_f16(tmem_base + g * 256, da + 2u * kk, db + 2u * kk, a.idesc_s, kk > 0 ? 1u : 0u);
          umma_commit(sfull_bar(g));
          pend[g].valid = 1; pend[g].stage = stage; pend[g].slot = slot; pend[g].last = slot == nslots - 1;
        }
        if (++stage == kFwdStages) { stage = 0; phase ^= 1u; }
      }
      if (pend[k & 1].valid) issue_pv(k & 1);                // the older of the two pending heads first
      if (pend[(k + 1) & 1].valid) issue_pv((k + 1) & 1);
    }
  } else {
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    RowCtx rc;
    rc.row = q * 32 + lane; rc.ws = rc.row >> 6; rc.i = rc.row & 63;
    rc.row_hi = rc.i >= 28; rc.col_hi = (kColHi >> min(rc.i, 63)) & 1ULL;
    const float* bias_row = bias_s + min(rc.i, kWt - 1) * kBiasPitch;
    const float sc2 = a.scale * kLog2e;
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 256 + 64 * rc.ws;
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 256 + kFwdTmemO + 32 * rc.ws;
    const uint32_t p_row = p_base + g * kTile + rc.row * 128;
    uint32_t n = 0;
    int k = 0;
    for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
      const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
      const int hp = unit - pair * nhp;
      const int nslots = min(2, a.heads - 2 * hp);
      row_unit(rc, a, pair);
      for (int slot = 0; slot < nslots; ++slot, ++k) {
        if ((k & 1) != g) continue;
        const int h = 2 * hp + slot;
        mbar_wait(sfull_bar(g), n & 1u);
        tc_fence_after();
        uint32_t r0[16], r1[16], r2[16], r3;
        tmem_ld16(t_s, r0);
        tmem_ld16(t_s + 16, r1);
        tmem_ld16(t_s + 32, r2);
        tmem_ld1(t_s + 48, r3);
        tmem_ld_wait();
        float s[kWt];
        scores_chunk<0, 16>(r0, bias_row, sc2, 0.f, rc.mask, s);
        scores_chunk<16, 16>(r1, bias_row, sc2, 0.f, rc.mask, s + 16);
        scores_chunk<32, 16>(r2, bias_row, sc2, 0.f, rc.mask, s + 32);
        scores_chunk<48, 1>(&r3, bias_row, sc2, 0.f, rc.mask, s + 48);
        float m = s[0];
#pragma unroll
        for (int j = 1; j < kWt; ++j) m = fmaxf(m, s[j]);
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < kWt; ++j) { s[j] = fast_exp2(s[j] - m); l += s[j]; }
        store_row_bf16(p_row, rc.row, s);            // un-normalised: 1 / l is applied to O
        fence_proxy_async_smem();                    // generic-proxy smem writes -> visible to the UMMA (async proxy)
        tc_fence_before();
        group_sync(g);
        if ((threadIdx.x & 127) == 64) mbar_arrive(pfull_bar(g));      // first thread of the group (threads 64.. / 192..)
        if (a.lse != nullptr && rc.valid) a.lse[rc.wm_row * a.heads + h] = (m + log2f(l)) * kLn2;
        mbar_wait(ofull_bar(g), n & 1u);
        tc_fence_after();
        store_row32(a.out + rc.gr * a.C + h * kHd, t_o, 1.0f / l, rc.valid);
        ++n;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
constexpr int kBwdStages = 2;
constexpr int kBwdStageBytes = 4 * kTile;                                  // q, k, v, dO tiles
constexpr int kBwdThreads = 128 + 2 * 128;                                 // producer, MMA, two dO loader warps, two groups
constexpr int kBwdTileBytes = kBwdStages * kBwdStageBytes + 4 * kTile;      // operand ring + (P, dS) tiles of each group
constexpr int kBwdSmem = kBwdTileBytes + kBiasBytes + kBarBytes + 1024;
static_assert(kBwdSmem <= 232448 && kFwdSmem <= 232448, "attention smem budget");
static_assert(2 * 128 * kWt * 4 <= kBwdStages * kBwdStageBytes, "rel-pos fold scratch fits in the operand ring");
// TMEM columns of group g (base 256 g): S at +0 and dP at +128 (128 columns each); once both have been read they are reused
// for dQ at +0 / +32 (window A / B), dK at +64 / +96, dV at +128 / +160

__global__ void __launch_bounds__(kBwdThreads, 1) window_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t pds_base = base + kBwdStages * kBwdStageBytes;          // group g: P tile at + 2 g kTile, dS tile after it
  float* bias_s = reinterpret_cast<float*>(base_ptr + kBwdTileBytes);
  const uint32_t bar_base = base + kBwdTileBytes + kBiasBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto dofull_bar = [&](int s) { return bar_base + 8u * (kBwdStages + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * kBwdStages + s); };
  auto sdp_bar = [&](int g) { return bar_base + 8u * (3 * kBwdStages + g); };           // S and dP complete
  auto pds_bar = [&](int g) { return bar_base + 8u * (3 * kBwdStages + 2 + g); };       // P and dS tiles written
  auto grads_bar = [&](int g) { return bar_base + 8u * (3 * kBwdStages + 4 + g); };     // dQ, dK, dV complete
  auto free_bar = [&](int g) { return bar_base + 8u * (3 * kBwdStages + 6 + g); };      // dQ, dK, dV read: the columns may be overwritten
  const uint32_t tmem_slot = bar_base + 8u * (3 * kBwdStages + 8);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kBwdTileBytes + kBiasBytes + 8 * (3 * kBwdStages + 8));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kBwdTileBytes / 16; i += kBwdThreads) st_shared_v4(base + 16u * i, 0u, 0u, 0u, 0u);
  build_bias_table(bias_s, a.pos);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int s = 0; s < kBwdStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(dofull_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(sdp_bar(g), 1); mbar_init(pds_bar(g), 1); mbar_init(grads_bar(g), 1); mbar_init(free_bar(g), 1); }
    mbar_fence_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_grid_sync();

  const int nhp = a.nhp;
  float acc[kWt];                              // groups: running sum of this row's dS over every task (rel-pos gradient)
#pragma unroll
  for (int j = 0; j < kWt; ++j) acc[j] = 0.f;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
        const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
        const int hp = unit - pair * nhp;
        mbar_wait_backoff(empty_bar(stage), phase ^ 1u, a.wait_ns);
        mbar_arrive_expect_tx(full_bar(stage), 6u * kBoxBytes);
        const uint32_t sb = base + stage * kBwdStageBytes;
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int w = 0; w < 2; ++w)
            tma_load_2d(sb + t * kTile + w * kWinOff, &tmap_qkv, full_bar(stage), t * a.C + hp * 64, (2 * pair + w) * kWt);
        if (++stage == kBwdStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      struct Pend { int valid, stage, slot, last; } pend[2] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
      uint32_t ng[2] = {0u, 0u};                 // gradient rounds issued per group (parity of pds)
      uint32_t ns[2] = {0u, 0u};                 // score rounds issued per group (parity of free)
      auto issue_grads = [&](int g) {
        const Pend pd = pend[g];
        mbar_wait_backoff(pds_bar(g), ng[g] & 1u, a.wait_ns);
        ++ng[g];
        tc_fence_after();
        const uint32_t sb = base + pd.stage * kBwdStageBytes;
        const uint32_t qs = sb, ks = sb + kTile, dos = sb + 3 * kTile;
        const uint32_t pt = pds_base + g * 2 * kTile, dt = pt + kTile;
        const uint32_t tg = tmem_base + g * 256;
        const uint64_t ds_k = make_sw128_desc(dt, 16, 1024);            // dS as the K-major A operand (contraction over keys)
        const uint64_t ds_mn = make_sw128_desc(dt, 8192, 1024);         // dS^T: MN-major, its two 64-row halves = two M chunks
        const uint64_t p_mn = make_sw128_desc(pt, 8192, 1024);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const uint32_t col = 64 * pd.slot;
          const uint64_t k_mn = make_sw128_desc(ks + w * kWinOff + col, 8192, 1024);
          const uint64_t q_mn = make_sw128_desc(qs + w * kWinOff + col, 8192, 1024);
          const uint64_t do_mn = make_sw128_desc(dos + w * kWinOff + col, 8192, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tg + 32 * w, ds_k + 2u * k, k_mn + 128u * k, a.idesc_kmn, k > 0 ? 1u : 0u);          // dQ
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tg + 64 + 32 * w, ds_mn + 128u * k, q_mn + 128u * k, a.idesc_mnmn, k > 0 ? 1u : 0u);  // dK
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tg + 128 + 32 * w, p_mn + 128u * k, do_mn + 128u * k, a.idesc_mnmn, k > 0 ? 1u : 0u);  // dV
        }
        umma_commit(grads_bar(g));
        if (pd.last) umma_commit(empty_bar(pd.stage));
        pend[g].valid = 0;
      };
      int k = 0;
      for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
        const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
        const int hp = unit - pair * nhp;
        const int nslots = min(2, a.heads - 2 * hp);
        for (int slot = 0; slot < nslots; ++slot, ++k) {
          const int g = k & 1;
          if (pend[g ^ 1].valid) issue_grads(g ^ 1);       // the previous head (other group): its P / dS are ready by now
          if (slot == 0) {
            mbar_wait_backoff(full_bar(stage), phase, a.wait_ns);
            mbar_wait_backoff(dofull_bar(stage), phase, a.wait_ns);
            tc_fence_after();
          }
          mbar_wait_backoff(free_bar(g), (ns[g] & 1u) ^ 1u, a.wait_ns);   // the group has read its previous dQ / dK / dV
          ++ns[g];
          tc_fence_after();
          const uint32_t sb = base + stage * kBwdStageBytes;
          const uint32_t col = 64 * slot;
          const uint64_t dq = make_sw128_desc(sb + col, 16, 1024), dk = make_sw128_desc(sb + kTile + col, 16, 1024);
          const uint64_t dv = make_sw128_desc(sb + 2 * kTile + col, 16, 1024), dd = make_sw128_desc(sb + 3 * kTile + col, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) umma_f16(tmem_base + g * 256, dq + 2u * kk, dk + 2u * kk, a.idesc_s, kk > 0 ? 1u : 0u);         // S = Q K^T
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) umma_f16(tmem_base + g * 256 + 128, dd + 2u * kk, dv + 2u * kk, a.idesc_s, kk > 0 ? 1u : 0u);   // dP = dO V^T
          umma_commit(sdp_bar(g));
          pend[g].valid = 1; pend[g].stage = stage; pend[g].slot = slot; pend[g].last = slot == nslots - 1;
        }
        if (++stage == kBwdStages) { stage = 0; phase ^= 1u; }
      }
      if (pend[k & 1].valid) issue_grads(k & 1);
      if (pend[(k + 1) & 1].valid) issue_grads((k + 1) & 1);
    }
  } else if (warp < 4) {
    // dO rows arrive in raster order: gather them into the swizzled tile with 16-B cp.async (8 lanes per 128-B row)
    const int lt = (warp - 2) * 32 + lane;
    const int chunk = lt & 7, rsub = lt >> 3;
    const int nww = a.W / kWs, nwh = a.H / kWs;
    const int off = a.shifted ? kWs / 2 : 0;
    int stage = 0; uint32_t phase = 0;
    int prev_stage = -1;
    for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
      const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
      const int hp = unit - pair * nhp;
      mbar_wait_backoff(empty_bar(stage), phase ^ 1u, a.wait_ns);
      const uint32_t dos = base + stage * kBwdStageBytes + 3 * kTile;
      const int col0 = hp * 64 + chunk * 8;
      if (col0 < a.C) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int win = 2 * pair + w;
          if (win < a.nwin) {
            const uint32_t wi = static_cast<uint32_t>(win);
            const uint32_t wrow = a.div_nww.div(wi);
            const int wx = static_cast<int>(wi - wrow * static_cast<uint32_t>(nww));
            const uint32_t img = a.div_nwh.div(wrow);
            const int wy = static_cast<int>(wrow - img * static_cast<uint32_t>(nwh));
#pragma unroll
            for (int pass = 0; pass < 7; ++pass) {
              const int t = pass * 8 + rsub;
              if (t < kWt) {
                const int r = t / kWs, c = t - r * kWs;
                int y = wy * kWs + r + off, x = wx * kWs + c + off;
                if (y >= a.H) y -= a.H;
                if (x >= a.W) x -= a.W;
                const long long gr = (static_cast<long long>(img) * a.H + y) * a.W + x;
                const int srow = w * 64 + t;
                cp_async16(dos + srow * 128 + ((chunk ^ (srow & 7)) << 4), a.dout + gr * a.C + col0);
              }
            }
          }
        }
      }
      cp_async_commit();
      if (prev_stage >= 0) {                     // the previous unit's rows have landed while this unit's were being issued
        cp_async_wait<1>();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(dofull_bar(prev_stage));
      }
      prev_stage = stage;
      if (++stage == kBwdStages) { stage = 0; phase ^= 1u; }
    }
    if (prev_stage >= 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(dofull_bar(prev_stage));
    }
  } else {
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    RowCtx rc;
    rc.row = q * 32 + lane; rc.ws = rc.row >> 6; rc.i = rc.row & 63;
    rc.row_hi = rc.i >= 28; rc.col_hi = (kColHi >> min(rc.i, 63)) & 1ULL;
    const float* bias_row = bias_s + min(rc.i, kWt - 1) * kBiasPitch;
    const float sc2 = a.scale * kLog2e;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 256;
    const uint32_t t_s = lane_addr + 64 * rc.ws, t_dp = lane_addr + 128 + 64 * rc.ws;
    const uint32_t p_row = pds_base + g * 2 * kTile + rc.row * 128, ds_row = p_row + kTile;
    const long long ld = 3LL * a.C;
    uint32_t n = 0;
    int k = 0;
    for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
      const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
      const int hp = unit - pair * nhp;
      const int nslots = min(2, a.heads - 2 * hp);
      row_unit(rc, a, pair);
      for (int slot = 0; slot < nslots; ++slot, ++k) {
        if ((k & 1) != g) continue;
        const int h = 2 * hp + slot;
        // padded rows: lse = +inf makes every probability exp2(-inf) = 0
        const float lse2 = rc.valid ? __ldg(a.lse + rc.wm_row * a.heads + h) * kLog2e : INFINITY;
        mbar_wait(sdp_bar(g), n & 1u);
        tc_fence_after();
        float p[kWt];
        float D = 0.f;
        {
          uint32_t r0[16], r1[16], r2[16], r3;
          tmem_ld16(t_s, r0);
          tmem_ld16(t_s + 16, r1);
          tmem_ld16(t_s + 32, r2);
          tmem_ld1(t_s + 48, r3);
          tmem_ld_wait();
          scores_chunk<0, 16>(r0, bias_row, sc2, lse2, rc.mask, p);
          scores_chunk<16, 16>(r1, bias_row, sc2, lse2, rc.mask, p + 16);
          scores_chunk<32, 16>(r2, bias_row, sc2, lse2, rc.mask, p + 32);
          scores_chunk<48, 1>(&r3, bias_row, sc2, lse2, rc.mask, p + 48);
#pragma unroll
          for (int j = 0; j < kWt; ++j) p[j] = fast_exp2(p[j]);
          store_row_bf16(p_row, rc.row, p);
        }
        {
          uint32_t d0[16], d1[16], d2[16], d3;
          tmem_ld16(t_dp, d0);
          tmem_ld16(t_dp + 16, d1);
          tmem_ld16(t_dp + 32, d2);
          tmem_ld1(t_dp + 48, d3);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            D = fmaf(p[j], __uint_as_float(d0[j]), D);
            D = fmaf(p[16 + j], __uint_as_float(d1[j]), D);
            D = fmaf(p[32 + j], __uint_as_float(d2[j]), D);
          }
          D = fmaf(p[48], __uint_as_float(d3), D);
          // dS = P o (dP - D), in place of P; its running sum is the rel-pos gradient of this row
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            p[j] *= __uint_as_float(d0[j]) - D;
            p[16 + j] *= __uint_as_float(d1[j]) - D;
            p[32 + j] *= __uint_as_float(d2[j]) - D;
          }
          p[48] *= __uint_as_float(d3) - D;
#pragma unroll
          for (int j = 0; j < kWt; ++j) acc[j] += p[j];
          store_row_bf16(ds_row, rc.row, p);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        group_sync(g);
        if ((threadIdx.x & 127) == 0) mbar_arrive(pds_bar(g));
        mbar_wait(grads_bar(g), n & 1u);
        tc_fence_after();
        bf16* dst = a.dqkv + rc.wm_row * ld + h * kHd;
        store_row32(dst, lane_addr + 32 * rc.ws, a.scale, rc.valid);                  // dQ (scaled: S = scale Q K^T + bias)
        store_row32(dst + a.C, lane_addr + 64 + 32 * rc.ws, a.scale, rc.valid);       // dK
        store_row32(dst + 2 * a.C, lane_addr + 128 + 32 * rc.ws, 1.0f, rc.valid);     // dV
        tc_fence_before();
        group_sync(g);
        if ((threadIdx.x & 127) == 0) mbar_arrive(free_bar(g));
        ++n;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  // Fold the per-row dS sums into the 13 x 13 bins, one partial row per CTA, fixed order (no atomics): bin (dr, dc) collects
  // every (query i, key j) with j's window row / column = i's + (dr, dc), over both groups and both window slots.
  float* fold = reinterpret_cast<float*>(base_ptr);                 // [2 groups][128 rows][49]: the operand ring is idle now
  if (warp >= 4) {
    const int g = (warp - 4) >> 2, row = (warp & 3) * 32 + lane;
#pragma unroll
    for (int j = 0; j < kWt; ++j) fold[(g * 128 + row) * kWt + j] = acc[j];
  }
  __syncthreads();
  if (threadIdx.x < kBins) {
    const int dr = static_cast<int>(threadIdx.x) / (2 * kWs - 1) - (kWs - 1), dc = static_cast<int>(threadIdx.x) % (2 * kWs - 1) - (kWs - 1);
    float sum = 0.f;
    for (int ri = max(0, -dr); ri < min(kWs, kWs - dr); ++ri)
      for (int ci = max(0, -dc); ci < min(kWs, kWs - dc); ++ci) {
        const int i = ri * kWs + ci, j = (ri + dr) * kWs + ci + dc;
#pragma unroll
        for (int src = 0; src < 4; ++src) sum += fold[((src >> 1) * 128 + (src & 1) * 64 + i) * kWt + j];
      }
    a.dpos_partial[1LL * blockIdx.x * kBins + threadIdx.x] = sum;
  }
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int fill_args(AttnArgs& a, int B, int H, int W, int C, int heads, int shifted) {
  B200_REQUIRE(B >= 0 && H > 0 && W > 0 && H % kWs == 0 && W % kWs == 0, "window_attn: H=%d W=%d must be multiples of 7", H, W);
  B200_REQUIRE(heads > 0 && C == heads * kHd, "window_attn: C=%d must equal heads(%d) * 32", C, heads);
  B200_REQUIRE(1LL * B * H * W < (1LL << 31) / 3, "window_attn: too many tokens");
  a.B = B; a.H = H; a.W = W; a.C = C; a.heads = heads; a.shifted = shifted;
  a.scale = 0.17677669529663687f;      // 32^-0.5
  a.nwin = B * (H / kWs) * (W / kWs);
  a.nhp = (heads + 1) / 2;
  a.units = (a.nwin + 1) / 2 * a.nhp;
  a.div_nhp = make_fastdiv(static_cast<uint32_t>(a.nhp));
  a.div_nww = make_fastdiv(static_cast<uint32_t>(W / kWs));
  a.div_nwh = make_fastdiv(static_cast<uint32_t>(H / kWs));
  a.idesc_s = attn_idesc(128, false, false);
  a.idesc_kmn = attn_idesc(32, false, true);
  a.idesc_mnmn = attn_idesc(32, true, true);
  a.wait_ns = static_cast<uint32_t>(b200_wait_ns());
  return B200_OK;
}

int qkv_tmap(CUtensorMap* map, const void* qkv, const AttnArgs& a) {
  B200_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "window_attn: qkv must be 16-B aligned");
  // [tokens (window-major), 3C] bf16; box = 64 columns (two heads) x the 49 rows of one window
  return gemm::encode_tmap_2d(map, true, qkv, 3ULL * a.C, static_cast<uint64_t>(a.nwin) * kWt, 3ULL * a.C, 64, kWt);
}

}  // namespace

extern "C" int b200_window_attn_fwd(const void* qkv, const float* pos, void* out, float* lse, int B, int H, int W, int C,
                                    int heads, int shifted, void* stream) {
  AttnArgs a{};
  int rc = fill_args(a, B, H, W, C, heads, shifted);
  if (rc) return rc;
  if (B == 0) return B200_OK;
  a.out = reinterpret_cast<bf16*>(out); a.lse = lse; a.pos = pos;
  B200_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "window_attn: out must be 16-B aligned");
  CUtensorMap tm;
  rc = qkv_tmap(&tm, qkv, a);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) { B200_CHECK_CUDA(cudaFuncSetAttribute(window_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem)); attr = true; }
  const int blocks = std::min(a.units, b200_num_sms());                 // persistent: one CTA per SM
  launch_pdl(window_attn_fwd_kernel, dim3(blocks), dim3(kFwdThreads), kFwdSmem, reinterpret_cast<cudaStream_t>(stream), tm, a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_window_attn_bwd_blocks(int B, int H, int W, int heads) {
  const long long units = (1LL * B * (H / kWs) * (W / kWs) + 1) / 2 * ((heads + 1) / 2);
  const long long cap = b200_num_sms();
  return static_cast<int>(std::max<long long>(1, std::min(units, cap)));
}

// floats of the `dpos_partial` scratch: one [169] partial row per CTA (padded to a 16-B multiple)
extern "C" long long b200_window_attn_bwd_scratch_floats(int blocks) { return (1LL * blocks * kBins + 3) / 4 * 4; }

extern "C" int b200_window_attn_bwd(const void* qkv, const float* pos, const float* lse, const void* dout,
                                    void* dqkv, float* dpos, float* dpos_partial, int accumulate_dpos, int B, int H, int W,
                                    int C, int heads, int shifted, void* stream) {
  AttnArgs a{};
  int rc = fill_args(a, B, H, W, C, heads, shifted);
  if (rc) return rc;
  if (B == 0) return B200_OK;
  a.pos = pos; a.lse = const_cast<float*>(lse); a.dout = reinterpret_cast<const bf16*>(dout); a.dqkv = reinterpret_cast<bf16*>(dqkv);
  a.dpos_partial = dpos_partial;
  B200_REQUIRE((reinterpret_cast<uintptr_t>(dout) & 15) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0, "window_attn: dout / dqkv must be 16-B aligned");
  CUtensorMap tm;
  rc = qkv_tmap(&tm, qkv, a);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) { B200_CHECK_CUDA(cudaFuncSetAttribute(window_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem)); attr = true; }
  const int blocks = b200_window_attn_bwd_blocks(B, H, W, heads);
  auto st = reinterpret_cast<cudaStream_t>(stream);
  launch_pdl(window_attn_bwd_kernel, dim3(blocks), dim3(kBwdThreads), kBwdSmem, st, tm, a);
  B200_LAUNCH_CHECK();
  // [blocks][169] partial rows -> the 13 x 13 table gradient, fixed order (recorded, not launched, inside a reduce batch)
  return reduce_or_defer(dpos_partial, &dpos, 1, kBins, blocks, accumulate_dpos, st, kBins);
}
