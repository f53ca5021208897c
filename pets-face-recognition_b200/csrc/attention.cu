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
// Warp roles (20 warps, one CTA per SM), a pipeline over the stream of (window pair, head) tasks:
//   TMA producer | issuer of the score MMAs | issuer of the second-stage MMAs (two threads, so neither waits behind the
//   other's barrier) | backward: two warps that gather the dO rows (raster order) with cp.async |
//   12 softmax warps - three per TMEM lane quarter, each thread owning one token row and 16 / 16 / 17 of its 49 key columns;
//   the parts of the row maximum / rowsum(P o dP) meet through smem - | 4 epilogue warps (TMEM -> bf16 -> global).
// Task n uses buffer n & 1 of everything that is double-buffered (P / dS tiles; forward: S and O accumulators), so the
// tensor core works on task n + 1 while the softmax warps are on task n and the epilogue warps on task n - 1.
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

// Three softmax warps share a TMEM lane quarter (the same 32 token rows); warp part T owns key columns [16 T, 16 T + 16)
// (+ column 48 for T = 2) of every row: 16 / 16 / 17 of the 49 keys
constexpr int kParts = 3;
constexpr int kPC = 17;                            // columns a thread holds at most
template <int T> struct Part { static constexpr int J0 = 16 * T; static constexpr int NC = T == 2 ? 17 : 16; };

// this thread's key columns [J0, J0 + NC) of its row: `taddr` = the row's first column
template <int T>
__device__ __forceinline__ void load_cols(uint32_t taddr, uint32_t (&v)[kPC]) {
  tmem_ld16(taddr + Part<T>::J0, reinterpret_cast<uint32_t(&)[16]>(v[0]));
  if (T == 2) tmem_ld1(taddr + 48, v[16]);
}

// accumulator columns -> scores in the log2 domain (scale * acc + bias - sub), shift-masked
template <int T>
__device__ __forceinline__ void scores(const uint32_t (&v)[kPC], const float* bias_row, float sc2, float sub, unsigned long long mask,
                                       float (&s)[kPC]) {
  constexpr int J0 = Part<T>::J0, NC = Part<T>::NC;
#pragma unroll
  for (int j4 = 0; j4 < (NC + 3) / 4; ++j4) {
    const float4 b = *reinterpret_cast<const float4*>(bias_row + J0 + 4 * j4);
    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (4 * j4 + e < NC) s[4 * j4 + e] = fmaf(__uint_as_float(v[4 * j4 + e]), sc2, bb[e] - sub);
  }
  if (mask != 0ULL) {                    // warp-uniform: a warp's 32 rows belong to one window
#pragma unroll
    for (int j = 0; j < NC; ++j)
      if ((mask >> (J0 + j)) & 1ULL) s[j] = -INFINITY;
  }
}

// the same with this thread's slice of its bias row held in registers for the whole kernel (a thread always serves the same
// token index, so the slice never changes)
template <int T>
__device__ __forceinline__ void scores_reg(const uint32_t (&v)[kPC], const float (&bias)[kPC], float sc2, unsigned long long mask, float (&s)[kPC]) {
  constexpr int J0 = Part<T>::J0, NC = Part<T>::NC;
#pragma unroll
  for (int j = 0; j < NC; ++j) s[j] = fmaf(__uint_as_float(v[j]), sc2, bias[j]);
  if (mask != 0ULL) {
#pragma unroll
    for (int j = 0; j < NC; ++j)
      if ((mask >> (J0 + j)) & 1ULL) s[j] = -INFINITY;
  }
}

// this thread's part of one row of a P / dS tile: bf16, 16-B chunks 2 T, 2 T + 1 (and chunk 6 = key 48 + zeros for T = 2) of the
// 128-B swizzled row; chunk 7 and keys 49..55 stay zero
template <int T>
__device__ __forceinline__ void store_part_row(uint32_t tile_row_addr, int row, const float (&v)[kPC]) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
    st_shared_v4(tile_row_addr + (((c + 2 * T) ^ (row & 7)) << 4), pack_bf16(v[8 * c], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]),
                 pack_bf16(v[8 * c + 4], v[8 * c + 5]), pack_bf16(v[8 * c + 6], v[8 * c + 7]));
  if (T == 2) st_shared_v4(tile_row_addr + ((6 ^ (row & 7)) << 4), pack_bf16(v[16], 0.f), 0u, 0u, 0u);
}

// 32 fp32 accumulator columns of this lane -> * f -> bf16 -> 64 B of global memory
__device__ __forceinline__ void load_row32(uint32_t taddr, uint32_t (&r0)[16], uint32_t (&r1)[16]) {
  tmem_ld16(taddr, r0);
  tmem_ld16(taddr + 16, r1);
}
__device__ __forceinline__ void store_row32(bf16* dst, const uint32_t (&r0)[16], const uint32_t (&r1)[16], float f) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
    st_global_v4(dst + 8 * c, pack_bf16(__uint_as_float(r0[8 * c]) * f, __uint_as_float(r0[8 * c + 1]) * f),
                 pack_bf16(__uint_as_float(r0[8 * c + 2]) * f, __uint_as_float(r0[8 * c + 3]) * f),
                 pack_bf16(__uint_as_float(r0[8 * c + 4]) * f, __uint_as_float(r0[8 * c + 5]) * f),
                 pack_bf16(__uint_as_float(r0[8 * c + 6]) * f, __uint_as_float(r0[8 * c + 7]) * f));
#pragma unroll
  for (int c = 0; c < 2; ++c)
    st_global_v4(dst + 16 + 8 * c, pack_bf16(__uint_as_float(r1[8 * c]) * f, __uint_as_float(r1[8 * c + 1]) * f),
                 pack_bf16(__uint_as_float(r1[8 * c + 2]) * f, __uint_as_float(r1[8 * c + 3]) * f),
                 pack_bf16(__uint_as_float(r1[8 * c + 4]) * f, __uint_as_float(r1[8 * c + 5]) * f),
                 pack_bf16(__uint_as_float(r1[8 * c + 6]) * f, __uint_as_float(r1[8 * c + 7]) * f));
}

constexpr int kThreads = 640;                      // 20 warps
constexpr int kSmWarp0 = 4;                        // softmax warps 4..15, epilogue warps 16..19
constexpr int kEpWarp0 = kSmWarp0 + 4 * kParts;
constexpr int kSmWarps = 4 * kParts;
// the softmax warps that share a TMEM lane quarter (and so the same 32 rows) meet on named barrier 1 + quarter
__device__ __forceinline__ void part_sync(int q) { asm volatile("bar.sync %0, 96;" ::"r"(1 + q) : "memory"); }
// one warp's arrival on a barrier that counts warps: every lane's prior work (smem writes + proxy fence, TMEM loads) first
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr int kFwdStages = 4;
constexpr int kFwdStageBytes = 3 * kTile;                                  // q, k, v tiles
constexpr int kFwdTileBytes = kFwdStages * kFwdStageBytes;                 // the operand ring (P lives in tensor memory)
constexpr int kXchBytes = 4096;                                             // row-maximum exchange between the column parts: xM [2][128][4]
constexpr int kFwdSmem = kFwdTileBytes + kBiasBytes + kXchBytes + kBarBytes + 1024;
// TMEM columns: S of buffer b at 128 b (128 columns); O of buffer b at 256 + 64 b (window A: +0, window B: +32); P of buffer
// b at 384 + 32 b: the A operand of O = P V read straight from tensor memory (bf16 pairs, key 2 c and 2 c + 1 in column c), so
// the probabilities never touch shared memory - the kernel is bound by shared-memory bandwidth (operand reads of the MMAs)
constexpr int kFwdTmemP = 384;
// row statistics for the epilogue warps, also in tensor memory: buffer b, columns 448 + 4 b .. + 3 = the three partial row
// sums (one per column part) and the row maximum - the softmax warps tcgen05.st them next to P, the epilogue tcgen05.ld's them
// next to O: no shared-memory hand-over
constexpr int kFwdTmemStat = 448;

template <int T>
__device__ __forceinline__ void fwd_softmax_task(const RowCtx& rc, const float (&bias)[kPC], float sc2, uint32_t t_s, uint32_t t_p, int q, int lane,
                                                 int b, uint32_t par, float* xM, uint32_t t_stat, uint32_t sfull, uint32_t sfree,
                                                 uint32_t ofree, uint32_t pfull) {
  constexpr int NC = Part<T>::NC;
  mbar_wait(sfull, par);
  tc_fence_after();
  uint32_t v[kPC];
  load_cols<T>(t_s, v);
  tmem_ld_wait();
  tc_fence_before();
  warp_arrive(sfree, lane);                      // the S buffer may be overwritten by the task after next
  float s[kPC];
  scores_reg<T>(v, bias, sc2, rc.mask, s);
  float m0 = s[0], m1 = s[1];
#pragma unroll
  for (int j = 2; j < NC; ++j) { if (j & 1) m1 = fmaxf(m1, s[j]); else m0 = fmaxf(m0, s[j]); }
  float m = fmaxf(m0, m1);
  float* xm = xM + (b * 128 + rc.row) * 4;
  xm[T] = m;
  part_sync(q);
  m = fmaxf(fmaxf(xm[0], xm[1]), xm[2]);
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    s[j] = fast_exp2(s[j] - m);
    if (j & 1) l1 += s[j]; else l0 += s[j];
  }
  // the epilogue warps have read O and the row statistics of the task before last (same buffers); its PV MMAs are done with
  // the P columns
  mbar_wait(ofree, par ^ 1u);
  tc_fence_after();
  tmem_st1(t_stat + T, __float_as_uint(l0 + l1));
  if (T == 0) tmem_st1(t_stat + 3, __float_as_uint(m));
  // un-normalised probabilities (1 / l is applied to O) -> bf16 pairs -> this row's P columns: keys 16 T .. 16 T + 15 are
  // columns 8 T .. 8 T + 7; part 2 also writes columns 24 .. 31 = key 48 and the zeros of the padded keys
  uint32_t pk[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) pk[c] = pack_bf16(s[2 * c], s[2 * c + 1]);
  tmem_st8(t_p + 8 * T, pk);
  if (T == 2) {
    uint32_t pz[8] = {pack_bf16(s[16], 0.f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    tmem_st8(t_p + 24, pz);
  }
  tmem_st_wait();
  tc_fence_before();
  warp_arrive(pfull, lane);
}

__global__ void __launch_bounds__(kThreads, 1) window_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(base_ptr + kFwdTileBytes);
  float* xM = reinterpret_cast<float*>(base_ptr + kFwdTileBytes + kBiasBytes);
  const uint32_t bar_base = base + kFwdTileBytes + kBiasBytes + kXchBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kFwdStages + s); };
  auto sfull_bar = [&](int b) { return bar_base + 8u * (2 * kFwdStages + b); };
  auto sfree_bar = [&](int b) { return bar_base + 8u * (2 * kFwdStages + 2 + b); };
  auto pfull_bar = [&](int b) { return bar_base + 8u * (2 * kFwdStages + 4 + b); };
  auto ofull_bar = [&](int b) { return bar_base + 8u * (2 * kFwdStages + 6 + b); };
  auto ofree_bar = [&](int b) { return bar_base + 8u * (2 * kFwdStages + 8 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kFwdStages + 10);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + kFwdTileBytes + kBiasBytes + kXchBytes + 8 * (2 * kFwdStages + 10));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows 49..63 / 113..127 of every tile are never written again: they must be zero
  for (int i = threadIdx.x; i < kFwdTileBytes / 16; i += kThreads) st_shared_v4(base + 16u * i, 0u, 0u, 0u, 0u);
  build_bias_table(bias_s, a.pos);      // pos is a parameter: not produced by the preceding kernel
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int s = 0; s < kFwdStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(sfull_bar(b), 1); mbar_init(sfree_bar(b), kSmWarps); mbar_init(pfull_bar(b), kSmWarps);
      mbar_init(ofull_bar(b), 1); mbar_init(ofree_bar(b), 4);
    }
    mbar_fence_init();
  }
  if (warp == 3) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
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
    if (lane == 0) {                             // S = Q K^T of task n into S buffer n & 1
      int stage = 0; uint32_t phase = 0, n = 0;
      for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
        const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
        const int nslots = min(2, a.heads - 2 * (unit - pair * nhp));
        mbar_wait_backoff(full_bar(stage), phase, a.wait_ns);
        for (int slot = 0; slot < nslots; ++slot, ++n) {
          const int b = n & 1;
          mbar_wait_backoff(sfree_bar(b), ((n >> 1) & 1u) ^ 1u, a.wait_ns);
          tc_fence_after();
          const uint32_t qs = base + stage * kFwdStageBytes, ks = qs + kTile;
          const uint64_t da = make_sw128_desc(qs + 64 * slot, 16, 1024);
          const uint64_t db = make_sw128_desc(ks + 64 * slot, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) umma_f16(tmem_base + b * 128, da + 2u * kk, db + 2u * kk, a.idesc_s, kk > 0 ? 1u : 0u);
          umma_commit(sfull_bar(b));
        }
        if (++stage == kFwdStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {                             // O = P V of task n into O buffer n & 1
      int stage = 0; uint32_t phase = 0, n = 0;
      for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
        const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
        const int nslots = min(2, a.heads - 2 * (unit - pair * nhp));
        mbar_wait_backoff(full_bar(stage), phase, a.wait_ns);
        for (int slot = 0; slot < nslots; ++slot, ++n) {
          const int b = n & 1;
          const uint32_t par = (n >> 1) & 1u;
          mbar_wait_backoff(pfull_bar(b), par, a.wait_ns);
          mbar_wait_backoff(ofree_bar(b), par ^ 1u, a.wait_ns);
          tc_fence_after();
          const uint32_t vs = base + stage * kFwdStageBytes + 2 * kTile;
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            // A = P from tensor memory (8 columns per 16 keys); B = V of window w, MN-major: contraction over its 64 key rows
            // (8-row groups 1024 B apart, 16 rows per MMA), N = the head's 32 columns inside the 64-column row
            const uint64_t db = make_sw128_desc(vs + w * kWinOff + 64 * slot, 8192, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ts(tmem_base + 256 + 64 * b + 32 * w, tmem_base + kFwdTmemP + 32 * b + 8 * k, db + 128u * k, a.idesc_kmn, k > 0 ? 1u : 0u);
          }
          umma_commit(ofull_bar(b));
          if (slot == nslots - 1) umma_commit(empty_bar(stage));     // every MMA that reads this unit's tiles has been issued
        }
        if (++stage == kFwdStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= kSmWarp0 && warp < kEpWarp0) {
    const int q = warp & 3, part = (warp - kSmWarp0) >> 2;
    RowCtx rc;
    rc.row = q * 32 + lane; rc.ws = rc.row >> 6; rc.i = rc.row & 63;
    rc.row_hi = rc.i >= 28; rc.col_hi = (kColHi >> min(rc.i, 63)) & 1ULL;
    const float sc2 = a.scale * kLog2e;
    float bias[kPC];
#pragma unroll
    for (int j = 0; j < kPC; ++j) bias[j] = bias_s[min(rc.i, kWt - 1) * kBiasPitch + min(16 * part + j, kBiasPitch - 1)];
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t lane_addr = lane_base + 64 * rc.ws;
    uint32_t n = 0;
    for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
      const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
      const int nslots = min(2, a.heads - 2 * (unit - pair * nhp));
      row_unit(rc, a, pair);
      for (int slot = 0; slot < nslots; ++slot, ++n) {
        const int b = n & 1;
        const uint32_t par = (n >> 1) & 1u;
        const uint32_t t_p = lane_base + kFwdTmemP + 32 * b, t_stat = lane_base + kFwdTmemStat + 4 * b;
        if (part == 0)
          fwd_softmax_task<0>(rc, bias, sc2, lane_addr + b * 128, t_p, q, lane, b, par, xM, t_stat, sfull_bar(b), sfree_bar(b),
                              ofree_bar(b), pfull_bar(b));
        else if (part == 1)
          fwd_softmax_task<1>(rc, bias, sc2, lane_addr + b * 128, t_p, q, lane, b, par, xM, t_stat, sfull_bar(b), sfree_bar(b),
                              ofree_bar(b), pfull_bar(b));
        else
          fwd_softmax_task<2>(rc, bias, sc2, lane_addr + b * 128, t_p, q, lane, b, par, xM, t_stat, sfull_bar(b), sfree_bar(b),
                              ofree_bar(b), pfull_bar(b));
      }
    }
  } else if (warp >= kEpWarp0) {
    const int q = warp & 3;
    RowCtx rc;
    rc.row = q * 32 + lane; rc.ws = rc.row >> 6; rc.i = rc.row & 63;
    rc.row_hi = false; rc.col_hi = false;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t lane_addr = lane_base + 256 + 32 * rc.ws;
    uint32_t n = 0;
    for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
      const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
      const int hp = unit - pair * nhp;
      const int nslots = min(2, a.heads - 2 * hp);
      row_unit(rc, a, pair);
      for (int slot = 0; slot < nslots; ++slot, ++n) {
        const int b = n & 1;
        const uint32_t par = (n >> 1) & 1u;
        const int h = 2 * hp + slot;
        mbar_wait(pfull_bar(b), par);            // the softmax warps' row statistics of this task are in tensor memory
        mbar_wait(ofull_bar(b), par);
        tc_fence_after();
        uint32_t r0[16], r1[16], st[4];
        load_row32(lane_addr + 64 * b, r0, r1);
        tmem_ld4(lane_base + kFwdTmemStat + 4 * b, st);
        tmem_ld_wait();
        const float l = (__uint_as_float(st[0]) + __uint_as_float(st[1])) + __uint_as_float(st[2]);
        const float m = __uint_as_float(st[3]);
        tc_fence_before();
        warp_arrive(ofree_bar(b), lane);
        if (rc.valid) {
          store_row32(a.out + rc.gr * a.C + h * kHd, r0, r1, 1.0f / l);
          if (a.lse != nullptr) a.lse[rc.wm_row * a.heads + h] = (m + log2f(l)) * kLn2;      // natural-log LSE
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
constexpr int kBwdStages = 2;
constexpr int kBwdStageBytes = 4 * kTile;                                  // q, k, v, dO tiles
constexpr int kBwdTileBytes = kBwdStages * kBwdStageBytes + 4 * kTile;      // operand ring + two (P, dS) tile pairs
constexpr int kBwdXchBytes = 4096;                                          // xD [2][128][4]
constexpr int kBwdSmem = kBwdTileBytes + kBiasBytes + kBwdXchBytes + kBarBytes + 1024;
static_assert(kBwdSmem <= 232448 && kFwdSmem <= 232448, "attention smem budget");
static_assert(128 * kWt * 4 <= kBwdStages * kBwdStageBytes, "rel-pos fold scratch fits in the operand ring");
// TMEM columns: S at 0, dP at 128 (128 columns each, single-buffered: the softmax warps pull them into registers at once);
// dQ at 256 / 288 (window A / B), dK at 320 / 352, dV at 384 / 416

template <int T>
__device__ __forceinline__ void bwd_softmax_task(const RowCtx& rc, const float* bias_row, float sc2, float lse2, uint32_t t_row, int q, int lane,
                                                 int b, uint32_t n, float* xD, uint32_t p_row, float (&acc)[kPC], uint32_t sdp, uint32_t sfree,
                                                 uint32_t gfull, uint32_t pds) {
  constexpr int NC = Part<T>::NC;
  mbar_wait(sdp, n & 1u);
  tc_fence_after();
  uint32_t v[kPC], d[kPC];
  load_cols<T>(t_row, v);
  load_cols<T>(t_row + 128, d);
  tmem_ld_wait();
  tc_fence_before();
  warp_arrive(sfree, lane);                      // S and dP are in registers: the next task's may be computed
  float p[kPC];
  scores<T>(v, bias_row, sc2, lse2, rc.mask, p);
  float D0 = 0.f, D1 = 0.f;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    p[j] = fast_exp2(p[j]);
    if (j & 1) D1 = fmaf(p[j], __uint_as_float(d[j]), D1); else D0 = fmaf(p[j], __uint_as_float(d[j]), D0);
  }
  float* xd = xD + (b * 128 + rc.row) * 4;
  xd[T] = D0 + D1;
  // the dK / dV MMAs of the task before last are done with this buffer's P / dS tiles
  mbar_wait(gfull, ((n >> 1) & 1u) ^ 1u);
  store_part_row<T>(p_row, rc.row, p);
  part_sync(q);
  const float D = (xd[0] + xd[1]) + xd[2];
  // dS = P o (dP - D), in place of P; its running sum is the rel-pos gradient of this row
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    p[j] *= __uint_as_float(d[j]) - D;
    acc[j] += p[j];
  }
  store_part_row<T>(p_row + kTile, rc.row, p);
  fence_proxy_async_smem();
  warp_arrive(pds, lane);
}

__global__ void __launch_bounds__(kThreads, 1) window_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t pds_base = base + kBwdStages * kBwdStageBytes;          // buffer b: P tile at + 2 b kTile, dS tile after it
  float* bias_s = reinterpret_cast<float*>(base_ptr + kBwdTileBytes);
  float* xD = reinterpret_cast<float*>(base_ptr + kBwdTileBytes + kBiasBytes);
  const uint32_t bar_base = base + kBwdTileBytes + kBiasBytes + kBwdXchBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto dofull_bar = [&](int s) { return bar_base + 8u * (kBwdStages + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * kBwdStages + s); };
  const uint32_t sdp_bar = bar_base + 8u * (3 * kBwdStages);                            // S and dP complete
  const uint32_t sfree_bar = bar_base + 8u * (3 * kBwdStages + 1);                      // ... and read
  auto pds_bar = [&](int b) { return bar_base + 8u * (3 * kBwdStages + 2 + b); };       // P and dS tiles of buffer b written
  auto gfull_bar = [&](int b) { return bar_base + 8u * (3 * kBwdStages + 4 + b); };     // dQ, dK, dV from buffer b complete
  const uint32_t gfree_bar = bar_base + 8u * (3 * kBwdStages + 6);                      // ... and read
  const uint32_t tmem_slot = bar_base + 8u * (3 * kBwdStages + 7);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + kBwdTileBytes + kBiasBytes + kBwdXchBytes + 8 * (3 * kBwdStages + 7));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kBwdTileBytes / 16; i += kThreads) st_shared_v4(base + 16u * i, 0u, 0u, 0u, 0u);
  build_bias_table(bias_s, a.pos);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int s = 0; s < kBwdStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(dofull_bar(s), 2); mbar_init(empty_bar(s), 1); }
    mbar_init(sdp_bar, 1); mbar_init(sfree_bar, kSmWarps); mbar_init(gfree_bar, 4);
    for (int b = 0; b < 2; ++b) { mbar_init(pds_bar(b), kSmWarps); mbar_init(gfull_bar(b), 1); }
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
  float acc[kPC];                              // softmax warps: running sum of this row's dS over every task (rel-pos gradient)
#pragma unroll
  for (int j = 0; j < kPC; ++j) acc[j] = 0.f;

  if (warp == 0) {
    if (lane == 0) {                             // S = Q K^T and dP = dO V^T of every task
      int stage = 0; uint32_t phase = 0, n = 0;
      for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
        const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
        const int nslots = min(2, a.heads - 2 * (unit - pair * nhp));
        mbar_wait_backoff(full_bar(stage), phase, a.wait_ns);
        mbar_wait_backoff(dofull_bar(stage), phase, a.wait_ns);
        for (int slot = 0; slot < nslots; ++slot, ++n) {
          mbar_wait_backoff(sfree_bar, (n & 1u) ^ 1u, a.wait_ns);
          tc_fence_after();
          const uint32_t sb = base + stage * kBwdStageBytes;
          const uint32_t col = 64 * slot;
          const uint64_t dq = make_sw128_desc(sb + col, 16, 1024), dk = make_sw128_desc(sb + kTile + col, 16, 1024);
          const uint64_t dv = make_sw128_desc(sb + 2 * kTile + col, 16, 1024), dd = make_sw128_desc(sb + 3 * kTile + col, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) umma_f16(tmem_base, dq + 2u * kk, dk + 2u * kk, a.idesc_s, kk > 0 ? 1u : 0u);         // S = Q K^T
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) umma_f16(tmem_base + 128, dd + 2u * kk, dv + 2u * kk, a.idesc_s, kk > 0 ? 1u : 0u);   // dP = dO V^T
          umma_commit(sdp_bar);
        }
        if (++stage == kBwdStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                             // dQ = dS K, dK = dS^T Q, dV = P^T dO of every task
      int stage = 0; uint32_t phase = 0, n = 0;
      for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
        const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
        const int nslots = min(2, a.heads - 2 * (unit - pair * nhp));
        mbar_wait_backoff(full_bar(stage), phase, a.wait_ns);
        mbar_wait_backoff(dofull_bar(stage), phase, a.wait_ns);
        for (int slot = 0; slot < nslots; ++slot, ++n) {
          const int b = n & 1;
          mbar_wait_backoff(pds_bar(b), (n >> 1) & 1u, a.wait_ns);
          mbar_wait_backoff(gfree_bar, (n & 1u) ^ 1u, a.wait_ns);      // the epilogue warps have read the previous dQ / dK / dV
          tc_fence_after();
          const uint32_t sb = base + stage * kBwdStageBytes;
          const uint32_t qs = sb, ks = sb + kTile, dos = sb + 3 * kTile;
          const uint32_t pt = pds_base + b * 2 * kTile, dt = pt + kTile;
          const uint64_t ds_k = make_sw128_desc(dt, 16, 1024);            // dS as the K-major A operand (contraction over keys)
          const uint64_t ds_mn = make_sw128_desc(dt, 8192, 1024);         // dS^T: MN-major, its two 64-row halves = two M chunks
          const uint64_t p_mn = make_sw128_desc(pt, 8192, 1024);
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const uint32_t col = 64 * slot;
            const uint64_t k_mn = make_sw128_desc(ks + w * kWinOff + col, 8192, 1024);
            const uint64_t q_mn = make_sw128_desc(qs + w * kWinOff + col, 8192, 1024);
            const uint64_t do_mn = make_sw128_desc(dos + w * kWinOff + col, 8192, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_base + 256 + 32 * w, ds_k + 2u * k, k_mn + 128u * k, a.idesc_kmn, k > 0 ? 1u : 0u);      // dQ
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_base + 320 + 32 * w, ds_mn + 128u * k, q_mn + 128u * k, a.idesc_mnmn, k > 0 ? 1u : 0u);  // dK
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_base + 384 + 32 * w, p_mn + 128u * k, do_mn + 128u * k, a.idesc_mnmn, k > 0 ? 1u : 0u);  // dV
          }
          umma_commit(gfull_bar(b));
          if (slot == nslots - 1) umma_commit(empty_bar(stage));
        }
        if (++stage == kBwdStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp < 4) {
    // q / k / v boxes by TMA (lane 0 of warp 2); the dO rows arrive in raster order: both warps gather them into the swizzled
    // tile with 16-B cp.async (8 lanes per 128-B row)
    const int lt = (warp - 2) * 32 + lane;
    const int chunk = lt & 7, rsub = lt >> 3;
    const int nww = a.W / kWs, nwh = a.H / kWs;
    const int off = a.shifted ? kWs / 2 : 0;
    int stage = 0; uint32_t phase = 0;
    int prev_stage = -1;
    for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
      const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
      const int hp = unit - pair * nhp;
      // the previous unit's rows are announced BEFORE this unit's slot is waited for: with two stages that wait ends only
      // when the unit before the previous one has been fully consumed, and the score MMAs of the previous unit must not
      // queue behind it
      if (prev_stage >= 0) {
        cp_async_wait<0>();
        fence_proxy_async_smem();
        warp_arrive(dofull_bar(prev_stage), lane);
      }
      mbar_wait_backoff(empty_bar(stage), phase ^ 1u, a.wait_ns);
      const uint32_t sb = base + stage * kBwdStageBytes;
      if (lt == 0) {
        mbar_arrive_expect_tx(full_bar(stage), 6u * kBoxBytes);
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int w = 0; w < 2; ++w)
            tma_load_2d(sb + t * kTile + w * kWinOff, &tmap_qkv, full_bar(stage), t * a.C + hp * 64, (2 * pair + w) * kWt);
      }
      const uint32_t dos = sb + 3 * kTile;
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
      prev_stage = stage;
      if (++stage == kBwdStages) { stage = 0; phase ^= 1u; }
    }
    if (prev_stage >= 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      warp_arrive(dofull_bar(prev_stage), lane);
    }
  } else if (warp < kEpWarp0) {
    const int q = warp & 3, part = (warp - kSmWarp0) >> 2;
    RowCtx rc;
    rc.row = q * 32 + lane; rc.ws = rc.row >> 6; rc.i = rc.row & 63;
    rc.row_hi = rc.i >= 28; rc.col_hi = (kColHi >> min(rc.i, 63)) & 1ULL;
    const float* bias_row = bias_s + min(rc.i, kWt - 1) * kBiasPitch;
    const float sc2 = a.scale * kLog2e;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 64 * rc.ws;
    uint32_t n = 0;
    // the row log-sum-exp of the NEXT task is fetched while the current one is computed (padded rows: +inf makes every
    // probability exp2(-inf) = 0)
    auto fetch_lse = [&](int unit, int slot) -> float {
      if (unit >= a.units) return INFINITY;
      const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
      const int hp = unit - pair * nhp;
      const int win = 2 * pair + rc.ws;
      if (win >= a.nwin || rc.i >= kWt) return INFINITY;
      return __ldg(a.lse + (static_cast<long long>(win) * kWt + rc.i) * a.heads + 2 * hp + slot) * kLog2e;
    };
    float lse_next = fetch_lse(blockIdx.x, 0);
    for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
      const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
      const int hp = unit - pair * nhp;
      const int nslots = min(2, a.heads - 2 * hp);
      row_unit(rc, a, pair);
      for (int slot = 0; slot < nslots; ++slot, ++n) {
        const int b = n & 1;
        const float lse2 = lse_next;
        lse_next = slot + 1 < nslots ? fetch_lse(unit, slot + 1) : fetch_lse(unit + static_cast<int>(gridDim.x), 0);
        const uint32_t p_row = pds_base + b * 2 * kTile + rc.row * 128;
        if (part == 0) bwd_softmax_task<0>(rc, bias_row, sc2, lse2, t_row, q, lane, b, n, xD, p_row, acc, sdp_bar, sfree_bar, gfull_bar(b), pds_bar(b));
        else if (part == 1) bwd_softmax_task<1>(rc, bias_row, sc2, lse2, t_row, q, lane, b, n, xD, p_row, acc, sdp_bar, sfree_bar, gfull_bar(b), pds_bar(b));
        else bwd_softmax_task<2>(rc, bias_row, sc2, lse2, t_row, q, lane, b, n, xD, p_row, acc, sdp_bar, sfree_bar, gfull_bar(b), pds_bar(b));
      }
    }
  } else {
    const int q = warp & 3;
    RowCtx rc;
    rc.row = q * 32 + lane; rc.ws = rc.row >> 6; rc.i = rc.row & 63;
    rc.row_hi = false; rc.col_hi = false;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 256 + 32 * rc.ws;
    const long long ld = 3LL * a.C;
    uint32_t n = 0;
    for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
      const int pair = static_cast<int>(a.div_nhp.div(static_cast<uint32_t>(unit)));
      const int hp = unit - pair * nhp;
      const int nslots = min(2, a.heads - 2 * hp);
      row_unit(rc, a, pair);
      for (int slot = 0; slot < nslots; ++slot, ++n) {
        const int b = n & 1;
        const int h = 2 * hp + slot;
        mbar_wait(gfull_bar(b), (n >> 1) & 1u);
        tc_fence_after();
        bf16* dst = a.dqkv + rc.wm_row * ld + h * kHd;
        uint32_t r0[16], r1[16], r2[16], r3[16];
        load_row32(lane_addr, r0, r1);                  // dQ
        load_row32(lane_addr + 64, r2, r3);             // dK
        tmem_ld_wait();
        if (rc.valid) {
          store_row32(dst, r0, r1, a.scale);            // scaled: S = scale Q K^T + bias
          store_row32(dst + a.C, r2, r3, a.scale);
        }
        load_row32(lane_addr + 128, r0, r1);            // dV
        tmem_ld_wait();
        tc_fence_before();
        warp_arrive(gfree_bar, lane);
        if (rc.valid) store_row32(dst + 2 * a.C, r0, r1, 1.0f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  // Fold the per-row dS sums into the 13 x 13 bins, one partial row per CTA, fixed order (no atomics): bin (dr, dc) collects
  // every (query i, key j) with j's window row / column = i's + (dr, dc), over both window slots.
  float* fold = reinterpret_cast<float*>(base_ptr);                 // [128 rows][49]: the operand ring is idle now
  if (warp >= kSmWarp0 && warp < kEpWarp0) {
    const int part = (warp - kSmWarp0) >> 2, row = (warp & 3) * 32 + lane;
#pragma unroll
    for (int j = 0; j < kPC; ++j)
      if (part == 2 || j < 16) fold[row * kWt + 16 * part + j] = acc[j];
  }
  __syncthreads();
  if (threadIdx.x < kBins) {
    const int dr = static_cast<int>(threadIdx.x) / (2 * kWs - 1) - (kWs - 1), dc = static_cast<int>(threadIdx.x) % (2 * kWs - 1) - (kWs - 1);
    float sum = 0.f;
    for (int ri = max(0, -dr); ri < min(kWs, kWs - dr); ++ri)
      for (int ci = max(0, -dc); ci < min(kWs, kWs - dc); ++ci) {
        const int i = ri * kWs + ci, j = (ri + dr) * kWs + ci + dc;
        sum += fold[i * kWt + j] + fold[(64 + i) * kWt + j];
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
  auto st = reinterpret_cast<cudaStream_t>(stream);
  // algorithmic work: q, k, v in, the output (and one LSE per row and head) out; 4 x 49 x 32 MACs per (token, head)
  const double tokens = 1.0 * B * H * W;
  const bool prof = b200_prof_kind_begin(st, B200_PROF_ATTN_FWD, tokens * heads * 4.0 * kWt * kHd * 2.0, tokens * C * 2.0 * 4.0 + (lse ? tokens * heads * 4.0 : 0.0));
  launch_pdl(window_attn_fwd_kernel, dim3(blocks), dim3(kThreads), kFwdSmem, st, tm, a);
  if (prof) b200_prof_kind_end(st);
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
  // algorithmic work: q, k, v, dO (and the LSE) in, dq, dk, dv out; five 49 x 49 x 32 products per (window, head)
  const double tokens = 1.0 * B * H * W;
  const bool prof = b200_prof_kind_begin(st, B200_PROF_ATTN_BWD, tokens * heads * 10.0 * kWt * kHd * 2.0, tokens * C * 2.0 * 7.0 + tokens * heads * 4.0);
  launch_pdl(window_attn_bwd_kernel, dim3(blocks), dim3(kThreads), kBwdSmem, st, tm, a);
  if (prof) b200_prof_kind_end(st);
  B200_LAUNCH_CHECK();
  // [blocks][169] partial rows -> the 13 x 13 table gradient, fixed order (recorded, not launched, inside a reduce batch)
  return reduce_or_defer(dpos_partial, &dpos, 1, kBins, blocks, accumulate_dpos, st, kBins);
}
