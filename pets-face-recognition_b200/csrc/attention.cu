// Fused (shifted-)window attention, forward and backward, for the berniwal-variant Swin block
// (reference models/swin.py:101-135).  One (window, head) problem is 49 tokens x 32 dims, padded to 64 x 32, and lives
// entirely in registers + a private smem slab of one warp (forward) or one warp pair (backward):
//
//   forward : S = Q K^T * scale + relpos + shift masks -> softmax -> O = P V            (P never leaves registers)
//   backward: recompute P from the saved row log-sum-exp; dV = P^T dO, dP = dO V^T,
//             dS = P o (dP - rowsum(P o dP)), dQ = scale dS K, dK = scale dS^T Q, dpos[bin] += dS
//
// The cyclic shift (roll -3 / +3) and the window partition are pure addressing here: token (r, c) of
// window (wy, wx) lives at pixel ((7 wy + r + off) mod H, (7 wx + c + off) mod W), off = 3 for shifted
// blocks, for the loads of q/k/v AND for the store of the result, so no rolled / rearranged copy of the
// activations is ever materialised.  q/k/v are read straight out of the [tokens, 3C] output of the qkv
// GEMM ([q|k|v] chunks, (head, dim) inside a chunk).
//
// 49-token problems are far below the 64/128-row tcgen05 atom and carry 3 % of the network's FLOPs, so
// the matmuls use warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate).  The kernels are bound by the latency of
// their gathered loads, so both are software-pipelined: the cp.async loads of a warp's NEXT task stream into the second
// half of a double-buffered slab while the current task is computed.  Tiles are 64-B rows with the 16-B chunk index
// XOR-swizzled by (row >> 1) & 3, which makes every ldmatrix phase conflict-free without padding.
#include "common.cuh"

#include "b200_fe.h"

namespace {

constexpr int kWs = 7;
constexpr int kWt = 49;
constexpr int kHd = 32;
constexpr int kRowB = 64;                          // bytes per tile row (32 bf16)
constexpr int kTileRows = 50;                      // 49 token rows + one all-zero row that stands in for rows 49..63
constexpr int kTileBytes = kTileRows * kRowB;      // 3200
constexpr int kPitch = 40;                         // bf16 elements per row of the 16-row output staging tile (80 B)
constexpr int kStageRowB = kPitch * 2;
constexpr int kStageBytes = 16 * kStageRowB;       // 1280
constexpr int kBins = 169;
constexpr int kBiasPitch = 72;                     // floats per row of the 64 x 64 bias table (conflict-free float2 reads)
constexpr int kBiasBytes = 64 * kBiasPitch * 4;    // 18432
constexpr float kLog2e = 1.4426950408889634f;
// bit j set <=> window column (j % 7) >= 4, i.e. token j lies in the wrapped part under the left/right mask
constexpr unsigned long long kColHi = 0x1C3870E1C3870ULL;

struct AttnArgs {
  const bf16* qkv;      // [B*H*W, 3C]
  bf16* out;            // fwd: [B*H*W, C] attention output (input of to_out)
  float* lse;           // [B*H*W, heads] row log-sum-exp (nullable in inference)
  const float* pos;     // [13*13] relative position table of this block
  const bf16* dout;     // bwd: grad wrt attention output [B*H*W, C]
  bf16* dqkv;           // bwd: [B*H*W, 3C]
  float* dpos_partial;  // bwd: [gridDim.x, 169]
  float* dslots;        // bwd: [gridDim.x * warps][64][32] per-lane dS accumulators (L2-resident scratch)
  int B, H, W, C, heads, shifted;
  float scale;
  FastDiv div_heads, div_nww, div_nwh;   // task -> (head, window column, window row, image)
  long long ntasks;
  int rev_tasks;
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float fast_exp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// byte offset of (tile row r, 16-B chunk c) inside a slab: rows >= 49 all read the shared zero row (row 49)
__device__ __forceinline__ uint32_t sw_off(int r, int c) {
  r = min(r, kWt);
  return static_cast<uint32_t>(r * kRowB + ((c ^ ((r >> 1) & 3)) << 4));
}
// per-lane ldmatrix offsets.  A-type: 16 rows of m-tile `mt`, k-step kk (16 dims).
__device__ __forceinline__ uint32_t off_a(int mt, int kk, int lane) { return sw_off(mt * 16 + (lane & 15), kk * 2 + (lane >> 4)); }
// B operand from an [n][k] slab, non-transposed: n-tile pair np (16 rows), k-step kk
__device__ __forceinline__ uint32_t off_b(int np, int kk, int lane) {
  return sw_off(np * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1));
}
// B operand from a [k][n] slab, transposed load: k-step kk (16 rows), n-tile pair np (16 dims)
__device__ __forceinline__ uint32_t off_bt(int kk, int np, int lane) {
  return sw_off(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 2 + (lane >> 4));
}

struct Task {
  int b, wy, wx, h;
  bool ul, lr;   // window gets the upper/lower resp. left/right shift mask (last row / last column of windows)
};

__device__ __forceinline__ Task decode_task(const AttnArgs& a, long long task) {
  const int nww = a.W / kWs, nwh = a.H / kWs;
  Task t;
  // rev_tasks: tasks are walked from the last image / window to the first (see b200_reverse_rows)
  const uint32_t tk = a.rev_tasks ? static_cast<uint32_t>(a.ntasks - 1 - task) : static_cast<uint32_t>(task);   // launcher: count fits 31 bits
  const uint32_t win = a.div_heads.div(tk);
  t.h = static_cast<int>(tk - win * static_cast<uint32_t>(a.heads));
  const uint32_t wrow = a.div_nww.div(win);
  t.wx = static_cast<int>(win - wrow * static_cast<uint32_t>(nww));
  const uint32_t img = a.div_nwh.div(wrow);
  t.wy = static_cast<int>(wrow - img * static_cast<uint32_t>(nwh));
  t.b = static_cast<int>(img);
  t.ul = a.shifted && (t.wy == nwh - 1);
  t.lr = a.shifted && (t.wx == nww - 1);
  return t;
}

// global token row of window token i (0..48)
__device__ __forceinline__ int token_row(const AttnArgs& a, const Task& t, int i) {
  const int off = a.shifted ? kWs / 2 : 0;
  const int r = i / kWs, c = i - r * kWs;
  int y = t.wy * kWs + r + off;
  int x = t.wx * kWs + c + off;
  if (y >= a.H) y -= a.H;
  if (x >= a.W) x -= a.W;
  return (t.b * a.H + y) * a.W + x;
}

// The 49 token rows of a task, two per lane (tokens lane and lane + 32), handed around by shuffles
struct Rows {
  int r0, r1;
  __device__ __forceinline__ void compute(const AttnArgs& a, const Task& t, int lane) {
    r0 = token_row(a, t, lane);
    r1 = lane + 32 < kWt ? token_row(a, t, lane + 32) : 0;
  }
  // row of token i; `hi` (i >= 32) must be warp-uniform
  __device__ __forceinline__ int get(int i, bool hi) const { return __shfl_sync(0xffffffffu, hi ? r1 : r0, i & 31); }
};

// 64 x 64 additive score table shared by every window and head of the block (models/swin.py:117-118):
// bias[i][j] = log2(e) * pos[r_j - r_i + 6][c_j - c_i + 6] (scores live in the log2 domain: one FFMA + EX2 per element);
// padded keys (j >= 49) -> -inf; padded queries (i >= 49) -> 0.
__device__ __forceinline__ void build_bias_table(float* bias_s, const float* __restrict__ pos) {
  for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
    const int i = idx >> 6, j = idx & 63;
    float v;
    if (j >= kWt) v = -INFINITY;
    else if (i >= kWt) v = 0.f;
    else {
      const int ri = i / kWs, ci = i - ri * kWs, rj = j / kWs, cj = j - rj * kWs;
      v = kLog2e * __ldg(pos + (rj - ri + kWs - 1) * (2 * kWs - 1) + (cj - ci + kWs - 1));
    }
    bias_s[i * kBiasPitch + j] = v;
  }
}

// shift masks (models/swin.py:49-62, 122-124) in closed form: -inf where query and key fall on different sides of
// the wrap boundary (window row >= 4 <=> token >= 28; window column >= 4 <=> bit of kColHi).

// Evaluated as bit vectors over a lane's accumulator columns (bit 2 n + e <-> key j = jbase + 8 n + e): the column
// sides are lane constants, a row's mask is one select per mask kind, and the per-element test is a constant-bit probe.
__device__ __forceinline__ void shift_col_bits(int jbase, int ntiles, uint32_t& col_ul, uint32_t& col_lr) {
  col_ul = 0u; col_lr = 0u;
  for (int n = 0; n < ntiles; ++n)
    for (int e = 0; e < 2; ++e) {
      const int j = jbase + n * 8 + e;
      col_ul |= (j >= 28 ? 1u : 0u) << (2 * n + e);
      col_lr |= static_cast<uint32_t>((kColHi >> j) & 1ULL) << (2 * n + e);
    }
}
__device__ __forceinline__ uint32_t shift_row_mask(const Task& t, int i, uint32_t col_ul, uint32_t col_lr) {
  uint32_t m = 0u;
  if (t.ul) m |= (i >= 28) ? ~col_ul : col_ul;
  if (t.lr) m |= ((kColHi >> i) & 1ULL) ? ~col_lr : col_lr;
  return m;
}

// stream one 49 x 32 bf16 tile (row i of the tile = window token i, 64 B = 4 lanes x 16 B) into a swizzled slab
__device__ __forceinline__ void load_tile_async(uint32_t slab, const bf16* base, long long ld, int col0, const Rows& rows, int lane) {
#pragma unroll
  for (int it = 0; it < 7; ++it) {
    const int row = it * 8 + (lane >> 2);
    const int gr = rows.get(row, it >= 4);
    if (row < kWt) cp_async16(slab + sw_off(row, lane & 3), base + 1LL * gr * ld + col0 + (lane & 3) * 8);
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// Two warps share one (window, head) task: its q / k / v tiles are loaded once, warp w computes query tiles 2w and 2w+1.
// Eight pairs (16 warps) per SM, each with a double-buffered slab.
constexpr int kFwdPairs = 8;
constexpr int kFwdThreads = kFwdPairs * 64;
constexpr int kFwdBufBytes = 3 * kTileBytes;                                     // q, k, v (q rows double as the O staging rows)
constexpr int kFwdPairBytes = 2 * kFwdBufBytes;
constexpr int kFwdSmem = kBiasBytes + kFwdPairs * kFwdPairBytes;

__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(kFwdThreads, 1) window_attn_fwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, w = warp & 1;
  const int g = lane >> 2, tq = lane & 3;
  float* bias_s = reinterpret_cast<float*>(smem);
  uint8_t* pb = smem + kBiasBytes + pair * kFwdPairBytes;
  const uint32_t pb_u = smem_u32(pb);
  const int bar_id = 1 + pair;

  build_bias_table(bias_s, a.pos);      // pos is a parameter: not produced by the preceding kernel
  for (int i = w * 32 + lane; i < kFwdPairBytes / 16; i += 64) reinterpret_cast<uint4*>(pb)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  pdl_grid_sync();

  const long long ntasks = 1LL * a.B * (a.H / kWs) * (a.W / kWs) * a.heads;
  uint32_t col_ul, col_lr;                 // this lane's key columns on the far side of the shift boundaries
  shift_col_bits((lane & 3) * 2, 8, col_ul, col_lr);
  const long long stride = 1LL * gridDim.x * kFwdPairs;
  const long long ld_qkv = 3LL * a.C;
  const float sc2 = a.scale * kLog2e;               // scores are kept in the log2 domain: one FFMA + EX2 per element
  // the two warps split the rows of each tile (8-row groups alternate)
  auto issue = [&](const Task& t, const Rows& rows, int buf) {
    const uint32_t base = pb_u + buf * kFwdBufBytes;
#pragma unroll
    for (int it = 0; it < 7; ++it) {
      const int row = it * 8 + (lane >> 2);
      const int gr = rows.get(row, it >= 4);
      if ((it & 1) == w && row < kWt) {
        const bf16* src = a.qkv + 1LL * gr * ld_qkv + t.h * kHd + (lane & 3) * 8;
        const uint32_t dst = base + sw_off(row, lane & 3);
        cp_async16(dst, src);
        cp_async16(dst + kTileBytes, src + a.C);
        cp_async16(dst + 2 * kTileBytes, src + 2 * a.C);
      }
    }
  };

  long long task = 1LL * blockIdx.x * kFwdPairs + pair;
  Task t{};
  Rows rows{};
  if (task < ntasks) {
    t = decode_task(a, task);
    rows.compute(a, t, lane);
    issue(t, rows, 0);
  }
  cp_async_commit();
  int buf = 0;
  for (; task < ntasks; task += stride) {
    cp_async_wait<0>();
    pair_sync(bar_id);          // both halves of this task's tiles have landed; the partner is done with the other buffer
    // the next task's tiles stream into the other buffer while this one is computed
    Task tn{};
    Rows rows_n{};
    if (task + stride < ntasks) {
      tn = decode_task(a, task + stride);
      rows_n.compute(a, tn, lane);
      issue(tn, rows_n, buf ^ 1);
    }
    cp_async_commit();
    const bool flagged = t.ul || t.lr;
    const uint32_t qs = pb_u + buf * kFwdBufBytes, ks = qs + kTileBytes, vs = ks + kTileBytes;
    uint8_t* qrows = pb + buf * kFwdBufBytes;

#pragma unroll 1
    for (int mi = 0; mi < 2; ++mi) {
      const int mt = 2 * w + mi;
      float s[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(qs + off_a(mt, kk, lane), a0, a1, a2, a3);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b0, b1, b2, b3;      // K as the B operand of S = Q K^T (n = key j, k = dim d)
          ldsm_x4(ks + off_b(np, kk, lane), b0, b1, b2, b3);
          mma16816(s[2 * np], a0, a1, a2, a3, b0, b1);
          mma16816(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
        }
      }
      const int i0 = mt * 16 + g, i1 = i0 + 8;
      const float* b0p = bias_s + i0 * kBiasPitch + tq * 2;
      const float* b1p = bias_s + i1 * kBiasPitch + tq * 2;
      float m0 = -INFINITY, m1 = -INFINITY;
      const uint32_t mk0 = flagged ? shift_row_mask(t, i0, col_ul, col_lr) : 0u;
      const uint32_t mk1 = flagged ? shift_row_mask(t, i1, col_ul, col_lr) : 0u;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const float2 b0 = *reinterpret_cast<const float2*>(b0p + n * 8);
        const float2 b1 = *reinterpret_cast<const float2*>(b1p + n * 8);
        s[n][0] = fmaf(s[n][0], sc2, b0.x);
        s[n][1] = fmaf(s[n][1], sc2, b0.y);
        s[n][2] = fmaf(s[n][2], sc2, b1.x);
        s[n][3] = fmaf(s[n][3], sc2, b1.y);
        if (flagged) {
          if (mk0 & (1u << (2 * n))) s[n][0] = -INFINITY;
          if (mk0 & (2u << (2 * n))) s[n][1] = -INFINITY;
          if (mk1 & (1u << (2 * n))) s[n][2] = -INFINITY;
          if (mk1 & (2u << (2 * n))) s[n][3] = -INFINITY;
        }
        m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
        m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      float l0 = 0.f, l1 = 0.f;
      uint32_t pf[8][2];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const float p0 = fast_exp2(s[n][0] - m0), p1 = fast_exp2(s[n][1] - m0);
        const float p2 = fast_exp2(s[n][2] - m1), p3 = fast_exp2(s[n][3] - m1);
        l0 += p0 + p1; l1 += p2 + p3;
        pf[n][0] = pack_bf16(p0, p1);
        pf[n][1] = pack_bf16(p2, p3);
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      float o[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b0, b1, b2, b3;        // V as the B operand of O = P V (k = key j, n = dim d): transposed ldmatrix
          ldsm_x4_t(vs + off_bt(kk, np, lane), b0, b1, b2, b3);
          mma16816(o[2 * np], pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1], b0, b1);
          mma16816(o[2 * np + 1], pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1], b2, b3);
        }
      const float r0 = 1.0f / l0, r1 = 1.0f / l1;
      __syncwarp();   // every lane has finished reading this m-tile's Q rows; reuse them as the O staging rows
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        if (i0 < kWt) *reinterpret_cast<uint32_t*>(qrows + sw_off(i0, n) + tq * 4) = pack_bf16(o[n][0] * r0, o[n][1] * r0);
        if (i1 < kWt) *reinterpret_cast<uint32_t*>(qrows + sw_off(i1, n) + tq * 4) = pack_bf16(o[n][2] * r1, o[n][3] * r1);
      }
      const int gr0 = rows.get(i0, w == 1), gr1 = rows.get(i1, w == 1);     // tokens >= 32 <=> the second warp's tiles
      if (a.lse != nullptr && tq == 0) {            // natural-log LSE of the scaled + biased scores
        if (i0 < kWt) a.lse[1LL * gr0 * a.heads + t.h] = (m0 + log2f(l0)) * 0.6931471805599453f;
        if (i1 < kWt) a.lse[1LL * gr1 * a.heads + t.h] = (m1 + log2f(l1)) * 0.6931471805599453f;
      }
    }
    __syncwarp();
#pragma unroll
    for (int it2 = 0; it2 < 4; ++it2) {       // this warp's 32 token rows
      const int row = (4 * w + it2) * 8 + (lane >> 2);
      const int gr = rows.get(row, w == 1);
      if (row < kWt) {
        const uint4 v = *reinterpret_cast<const uint4*>(qrows + sw_off(row, lane & 3));
        *reinterpret_cast<uint4*>(a.out + 1LL * gr * a.C + t.h * kHd + (lane & 3) * 8) = v;
      }
    }
    t = tn;
    rows = rows_n;
    buf ^= 1;
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// Two warps share one (window, head) task; warp w owns keys [32 w, 32 w + 32).  Per 16-query tile each warp forms its
// 16 x 32 slice of S and dP once (query-major), turns it into P and dS, and uses those fragments three ways:
//   dQ_part = dS K_w          (A = dS straight from the accumulator registers; the two halves are summed through smem)
//   dV_w   += P^T dO,  dK_w += dS^T Q      (A = the 8 x 8 blocks transposed in registers by movmatrix)
// so no score is recomputed in the key-major orientation and dK / dV never leave registers until the task ends.
// rowsum(P o dP) is formed from the same fragments (the halves exchange their partial sums through smem), so the saved
// attention output is not read at all.
constexpr int kPairs = 6;                          // tasks in flight per CTA (one CTA of 12 warps per SM)
constexpr int kBwdThreads = kPairs * 64;
constexpr int kSlots = 4 * 4 * 4;                  // per-lane dS accumulators: [m-tile][n-tile of the warp's 32 keys][fragment element]
constexpr int kXBytes = 16 * 32 * 4;               // one warp's dQ partial of a 16-query tile, fp32
constexpr int kBwdBufBytes = 4 * kTileBytes + 64 * 4;                            // q, k, v, dO tiles + the 64 row LSEs
// per pair: two buffers, two dQ hand-over slabs, a staging tile per warp, rowsum partials [2][2 warps][16], 2 mbarriers
constexpr int kPairBytes = 2 * kBwdBufBytes + 2 * kXBytes + 2 * kStageBytes + 2 * 2 * 16 * 4 + 16;
constexpr int kBwdSmem = kBiasBytes + 704 + kPairs * kPairBytes;
static_assert(kPairBytes % 16 == 0 && kBwdSmem <= 227 * 1024, "attention backward smem layout");

__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ void red_add_f32x4(float* p, float v0, float v1, float v2, float v3) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
}

__global__ void __launch_bounds__(kBwdThreads, 1) window_attn_bwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, w = warp & 1;
  const int g = lane >> 2, tq = lane & 3;
  float* bias_s = reinterpret_cast<float*>(smem);
  uint8_t* pb = smem + kBiasBytes + 704 + pair * kPairBytes;
  const uint32_t pb_u = smem_u32(pb);
  float* xbuf = reinterpret_cast<float*>(pb + 2 * kBwdBufBytes);                                // [2][16][32]
  bf16* stage = reinterpret_cast<bf16*>(pb + 2 * kBwdBufBytes + 2 * kXBytes + w * kStageBytes);
  float* dpart_all = reinterpret_cast<float*>(pb + 2 * kBwdBufBytes + 2 * kXBytes + 2 * kStageBytes);  // [2][2 warps][16]
  const uint32_t xbar = pb_u + 2 * kBwdBufBytes + 2 * kXBytes + 2 * kStageBytes + 2 * 2 * 16 * 4;      // 2 mbarriers
  float* slots = a.dslots + (1LL * blockIdx.x * (2 * kPairs) + warp) * (kSlots * 32) + lane * 4;   // [slot / 4][lane][4]
  const int bar_id = 1 + pair;                              // named barrier of this pair

  build_bias_table(bias_s, a.pos);
  for (int i = w * 32 + lane; i < kPairBytes / 16; i += 64) reinterpret_cast<uint4*>(pb)[i] = make_uint4(0, 0, 0, 0);
#pragma unroll 4
  for (int s4 = 0; s4 < kSlots / 4; ++s4) __stcg(reinterpret_cast<float4*>(slots + s4 * 128), make_float4(0.f, 0.f, 0.f, 0.f));
  __syncthreads();
  if (w == 0 && lane == 0) {
    mbar_init(xbar, 1);
    mbar_init(xbar + 8, 1);
    mbar_fence_init();
  }
  // padded query rows: exp2(x - inf) = 0
  if (w == 1 && lane >= kWt - 32) {
    *reinterpret_cast<float*>(pb + 4 * kTileBytes + (32 + lane) * 4) = INFINITY;
    *reinterpret_cast<float*>(pb + kBwdBufBytes + 4 * kTileBytes + (32 + lane) * 4) = INFINITY;
  }
  __syncthreads();
  pdl_grid_sync();

  const long long ntasks = 1LL * a.B * (a.H / kWs) * (a.W / kWs) * a.heads;
  uint32_t col_ul, col_lr;                 // this lane's key columns on the far side of the shift boundaries
  shift_col_bits(w * 32 + tq * 2, 4, col_ul, col_lr);
  const long long stride = 1LL * gridDim.x * kPairs;
  const long long ld_qkv = 3LL * a.C;
  const float sc2 = a.scale * kLog2e;
  // warp 0 streams q and k, warp 1 v and dO; each warp the LSE of "its" 32 query rows
  auto issue = [&](const Task& t, const Rows& rows, int buf) {
    const uint32_t base = pb_u + buf * kBwdBufBytes;
    if (w == 0) {
      load_tile_async(base, a.qkv, ld_qkv, t.h * kHd, rows, lane);
      load_tile_async(base + kTileBytes, a.qkv, ld_qkv, a.C + t.h * kHd, rows, lane);
    } else {
      load_tile_async(base + 2 * kTileBytes, a.qkv, ld_qkv, 2 * a.C + t.h * kHd, rows, lane);
      load_tile_async(base + 3 * kTileBytes, a.dout, a.C, t.h * kHd, rows, lane);
    }
    const int i = w * 32 + lane;
    if (i < kWt) cp_async4(base + 4 * kTileBytes + i * 4, a.lse + 1LL * (w ? rows.r1 : rows.r0) * a.heads + t.h);
  };

  long long task = 1LL * blockIdx.x * kPairs + pair;
  Task t{};
  Rows rows{};
  if (task < ntasks) {
    t = decode_task(a, task);
    rows.compute(a, t, lane);
    issue(t, rows, 0);
  }
  cp_async_commit();
  int buf = 0;
  for (; task < ntasks; task += stride) {
    cp_async_wait<0>();
    pair_sync(bar_id);          // both halves of this task's tiles have landed; the partner is done with the other buffer
    Task tn{};
    Rows rows_n{};
    if (task + stride < ntasks) {
      tn = decode_task(a, task + stride);
      rows_n.compute(a, tn, lane);
      issue(tn, rows_n, buf ^ 1);
    }
    cp_async_commit();
    const bool flagged = t.ul || t.lr;
    const uint32_t qs = pb_u + buf * kBwdBufBytes, ks = qs + kTileBytes, vs = ks + kTileBytes, dos = vs + kTileBytes;
    const float* lse_s = reinterpret_cast<const float*>(pb + buf * kBwdBufBytes + 4 * kTileBytes);

    float dv[2][4][4], dk[2][4][4];       // this warp's 32 keys x 32 dims, accumulated over the four query tiles
#pragma unroll
    for (int jt = 0; jt < 2; ++jt)
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        dv[jt][n][0] = dv[jt][n][1] = dv[jt][n][2] = dv[jt][n][3] = 0.f;
        dk[jt][n][0] = dk[jt][n][1] = dk[jt][n][2] = dk[jt][n][3] = 0.f;
      }

#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      float s[4][4], dp[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t q0, q1, q2, q3, d0, d1, d2, d3;
        const uint32_t aoff = off_a(mt, kk, lane);
        ldsm_x4(qs + aoff, q0, q1, q2, q3);
        ldsm_x4(dos + aoff, d0, d1, d2, d3);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const uint32_t boff = off_b(2 * w + np, kk, lane);
          ldsm_x4(ks + boff, b0, b1, b2, b3);   // B = K (n = key j, k = d)
          mma16816(s[2 * np], q0, q1, q2, q3, b0, b1);
          mma16816(s[2 * np + 1], q0, q1, q2, q3, b2, b3);
          ldsm_x4(vs + boff, b0, b1, b2, b3);   // B = V (n = key j, k = d)
          mma16816(dp[2 * np], d0, d1, d2, d3, b0, b1);
          mma16816(dp[2 * np + 1], d0, d1, d2, d3, b2, b3);
        }
      }
      const int i0 = mt * 16 + g, i1 = i0 + 8;
      const int jw = w * 32 + tq * 2;
      const float* b0p = bias_s + i0 * kBiasPitch + jw;
      const float* b1p = bias_s + i1 * kBiasPitch + jw;
      const float l0 = lse_s[i0] * kLog2e, l1 = lse_s[i1] * kLog2e;
      // P (in place of S) and this half's share of D_i = sum_j P_ij dP_ij
      float D0 = 0.f, D1 = 0.f;
      const uint32_t mk0 = flagged ? shift_row_mask(t, i0, col_ul, col_lr) : 0u;
      const uint32_t mk1 = flagged ? shift_row_mask(t, i1, col_ul, col_lr) : 0u;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const float2 b0 = *reinterpret_cast<const float2*>(b0p + n * 8);
        const float2 b1 = *reinterpret_cast<const float2*>(b1p + n * 8);
        float sc[4];
        sc[0] = fmaf(s[n][0], sc2, b0.x - l0); sc[1] = fmaf(s[n][1], sc2, b0.y - l0);
        sc[2] = fmaf(s[n][2], sc2, b1.x - l1); sc[3] = fmaf(s[n][3], sc2, b1.y - l1);
        if (flagged) {
          if (mk0 & (1u << (2 * n))) sc[0] = -INFINITY;
          if (mk0 & (2u << (2 * n))) sc[1] = -INFINITY;
          if (mk1 & (1u << (2 * n))) sc[2] = -INFINITY;
          if (mk1 & (2u << (2 * n))) sc[3] = -INFINITY;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) s[n][e] = fast_exp2(sc[e]);
        D0 = fmaf(s[n][0], dp[n][0], fmaf(s[n][1], dp[n][1], D0));
        D1 = fmaf(s[n][2], dp[n][2], fmaf(s[n][3], dp[n][3], D1));
      }
      D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
      D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
      float* dpart = dpart_all + (mt & 1) * 32;        // double-buffered: the partner may still be reading the previous tile's
      if (tq == 0) { dpart[w * 16 + g] = D0; dpart[w * 16 + g + 8] = D1; }
      pair_sync(bar_id);
      D0 += dpart[(w ^ 1) * 16 + g];
      D1 += dpart[(w ^ 1) * 16 + g + 8];
      // dS; its per-lane running sums (for the rel-pos gradient) live in an L2-resident scratch and are updated by
      // fire-and-forget reductions: every address is private to one lane, so the order of the additions is fixed
      float* sl = slots + mt * (4 * 128);
      uint32_t pf[4][2], df[4][2];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        float ds[4];
        ds[0] = s[n][0] * (dp[n][0] - D0); ds[1] = s[n][1] * (dp[n][1] - D0);       // unscaled: `scale` is applied once to dQ / dK
        ds[2] = s[n][2] * (dp[n][2] - D1); ds[3] = s[n][3] * (dp[n][3] - D1);
        red_add_f32x4(sl + n * 128, ds[0], ds[1], ds[2], ds[3]);
        pf[n][0] = pack_bf16(s[n][0], s[n][1]); pf[n][1] = pack_bf16(s[n][2], s[n][3]);
        df[n][0] = pack_bf16(ds[0], ds[1]); df[n][1] = pack_bf16(ds[2], ds[3]);
      }
      // dQ (this warp's 32 keys): contraction over keys j
      float dq[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) { dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(ks + off_bt(2 * w + kk, np, lane), b0, b1, b2, b3);    // B = K (k = j, n = d)
          mma16816(dq[2 * np], df[2 * kk][0], df[2 * kk][1], df[2 * kk + 1][0], df[2 * kk + 1][1], b0, b1);
          mma16816(dq[2 * np + 1], df[2 * kk][0], df[2 * kk][1], df[2 * kk + 1][0], df[2 * kk + 1][1], b2, b3);
        }
      // the half that does not finish this tile's dQ hands its partial over right away (double-buffered slab + mbarrier)
      const bool finisher = w == (mt & 1);
      float* xb = xbuf + (mt & 1) * (16 * 32) + lane;
      if (!finisher) {
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) xb[(n * 4 + e) * 32] = dq[n][e];
        __syncwarp();
        if (lane == 0) mbar_arrive(xbar + 8 * (mt & 1));
      }
      // P^T and dS^T as A operands (rows = keys, k = the 16 queries of this tile): transpose the 8 x 8 blocks in registers
      uint32_t pt[4][2], dt[4][2];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        pt[n][0] = movm_t(pf[n][0]); pt[n][1] = movm_t(pf[n][1]);
        dt[n][0] = movm_t(df[n][0]); dt[n][1] = movm_t(df[n][1]);
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const uint32_t boff = off_bt(mt, np, lane);
        ldsm_x4_t(dos + boff, b0, b1, b2, b3);     // B = dO (k = i, n = d)
#pragma unroll
        for (int jt = 0; jt < 2; ++jt) {
          mma16816(dv[jt][2 * np], pt[2 * jt][0], pt[2 * jt + 1][0], pt[2 * jt][1], pt[2 * jt + 1][1], b0, b1);
          mma16816(dv[jt][2 * np + 1], pt[2 * jt][0], pt[2 * jt + 1][0], pt[2 * jt][1], pt[2 * jt + 1][1], b2, b3);
        }
        ldsm_x4_t(qs + boff, b0, b1, b2, b3);      // B = Q (k = i, n = d)
#pragma unroll
        for (int jt = 0; jt < 2; ++jt) {
          mma16816(dk[jt][2 * np], dt[2 * jt][0], dt[2 * jt + 1][0], dt[2 * jt][1], dt[2 * jt + 1][1], b0, b1);
          mma16816(dk[jt][2 * np + 1], dt[2 * jt][0], dt[2 * jt + 1][0], dt[2 * jt][1], dt[2 * jt + 1][1], b2, b3);
        }
      }
      const int gri = rows.get(mt * 16 + (lane >> 2), mt >= 2), gri8 = rows.get(mt * 16 + 8 + (lane >> 2), mt >= 2);
      if (finisher) {
        mbar_wait(xbar + 8 * (mt & 1), static_cast<uint32_t>(mt >> 1));     // each barrier completes twice per task
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) dq[n][e] += xb[(n * 4 + e) * 32];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          *reinterpret_cast<uint32_t*>(stage + g * kPitch + n * 8 + tq * 2) = pack_bf16(dq[n][0] * a.scale, dq[n][1] * a.scale);
          *reinterpret_cast<uint32_t*>(stage + (g + 8) * kPitch + n * 8 + tq * 2) = pack_bf16(dq[n][2] * a.scale, dq[n][3] * a.scale);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int r = it * 8 + (lane >> 2);
          if (mt * 16 + r < kWt) {
            const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(stage) + r * kStageRowB + (lane & 3) * 16);
            *reinterpret_cast<uint4*>(a.dqkv + 1LL * (it ? gri8 : gri) * ld_qkv + t.h * kHd + (lane & 3) * 8) = v;
          }
        }
        __syncwarp();
      }
    }
    // dK (scaled) and dV rows of this warp's keys: stage 16 rows at a time, 64-B row segments out
#pragma unroll
    for (int jt = 0; jt < 2; ++jt) {
      const int grj = rows.get(jt * 16 + (lane >> 2), w == 1), grj8 = rows.get(jt * 16 + 8 + (lane >> 2), w == 1);
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const float(&src)[4] = which == 0 ? dk[jt][n] : dv[jt][n];
          const float f = which == 0 ? a.scale : 1.0f;
          *reinterpret_cast<uint32_t*>(stage + g * kPitch + n * 8 + tq * 2) = pack_bf16(src[0] * f, src[1] * f);
          *reinterpret_cast<uint32_t*>(stage + (g + 8) * kPitch + n * 8 + tq * 2) = pack_bf16(src[2] * f, src[3] * f);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int r = it * 8 + (lane >> 2);
          const int j = w * 32 + jt * 16 + r;
          if (j < kWt) {
            const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(stage) + r * kStageRowB + (lane & 3) * 16);
            *reinterpret_cast<uint4*>(a.dqkv + 1LL * (it ? grj8 : grj) * ld_qkv + (which == 0 ? 1 : 2) * a.C + t.h * kHd + (lane & 3) * 8) = v;
          }
        }
      }
    }
    t = tn;
    rows = rows_n;
    buf ^= 1;
  }
  cp_async_wait<0>();
  // Fold the per-lane accumulators into the 13 x 13 bins: one partial row per CTA.  Thread b gathers bin b straight from
  // the L2-resident slots of the CTA's 12 warps in a fixed order (no atomics): bin (dr, dc) collects every (query i, key j)
  // with j's window row / column = i's + (dr, dc), and (i, j) lives in warp-half w = j / 32, slot group (i / 16) * 4 +
  // (j % 32) / 8, lane (i % 8) * 4 + (j % 8) / 2, element ((i % 16) / 8) * 2 + (j % 2) of the mma accumulator layout.
  __threadfence();
  __syncthreads();
  if (threadIdx.x < kBins) {
    const int dr = static_cast<int>(threadIdx.x) / (2 * kWs - 1) - (kWs - 1), dc = static_cast<int>(threadIdx.x) % (2 * kWs - 1) - (kWs - 1);
    const float* cta_slots = a.dslots + 1LL * blockIdx.x * (2 * kPairs) * (kSlots * 32);
    float acc = 0.f;
    for (int ri = max(0, -dr); ri < min(kWs, kWs - dr); ++ri)
      for (int ci = max(0, -dc); ci < min(kWs, kWs - dc); ++ci) {
        const int i = ri * kWs + ci, j = (ri + dr) * kWs + ci + dc;
        const int off = (j >> 5) * (kSlots * 32) + (((i >> 4) * 4 + ((j & 31) >> 3)) * 32 + (i & 7) * 4 + ((j & 7) >> 1)) * 4 +
                        ((i & 15) >> 3) * 2 + (j & 1);
#pragma unroll
        for (int pr = 0; pr < kPairs; ++pr) acc += __ldcg(cta_slots + pr * (2 * kSlots * 32) + off);
      }
    a.dpos_partial[1LL * blockIdx.x * kBins + threadIdx.x] = acc;
  }
}


int check_shape(int B, int H, int W, int C, int heads) {
  B200_REQUIRE(B >= 0 && H > 0 && W > 0 && H % kWs == 0 && W % kWs == 0, "window_attn: H=%d W=%d must be multiples of 7", H, W);
  B200_REQUIRE(heads > 0 && C == heads * kHd, "window_attn: C=%d must equal heads(%d) * 32", C, heads);
  return B200_OK;
}

}  // namespace

extern "C" int b200_window_attn_fwd(const void* qkv, const float* pos, void* out, float* lse, int B, int H, int W, int C,
                                    int heads, int shifted, void* stream) {
  int rc = check_shape(B, H, W, C, heads);
  if (rc) return rc;
  if (B == 0) return B200_OK;
  AttnArgs a{};
  a.qkv = reinterpret_cast<const bf16*>(qkv); a.out = reinterpret_cast<bf16*>(out); a.lse = lse; a.pos = pos;
  a.B = B; a.H = H; a.W = W; a.C = C; a.heads = heads; a.shifted = shifted; a.scale = 0.17677669529663687f;  // 32^-0.5
  a.div_heads = make_fastdiv(static_cast<uint32_t>(heads)); a.div_nww = make_fastdiv(static_cast<uint32_t>(W / kWs)); a.div_nwh = make_fastdiv(static_cast<uint32_t>(H / kWs));
  a.ntasks = 1LL * B * (H / kWs) * (W / kWs) * heads; a.rev_tasks = b200_reverse_rows();
  static bool attr = false;
  if (!attr) { B200_CHECK_CUDA(cudaFuncSetAttribute(window_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem)); attr = true; }
  const long long ntasks = 1LL * B * (H / kWs) * (W / kWs) * heads;
  B200_REQUIRE(ntasks < (1LL << 31), "window_attn: %lld (window, head) tasks exceed 31 bits", ntasks);
  long long blocks = (ntasks + kFwdPairs - 1) / kFwdPairs;
  const long long cap = b200_num_sms();                   // persistent: one CTA per SM
  if (blocks > cap) blocks = cap;
  launch_pdl(window_attn_fwd_kernel, dim3(static_cast<unsigned>(blocks)), dim3(kFwdThreads), kFwdSmem, reinterpret_cast<cudaStream_t>(stream), a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_window_attn_bwd_blocks(int B, int H, int W, int heads) {
  const long long ntasks = 1LL * B * (H / kWs) * (W / kWs) * heads;
  long long blocks = (ntasks + kPairs - 1) / kPairs;
  const long long cap = 1LL * b200_num_sms();
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : static_cast<int>(blocks);
}

// floats per CTA of the scratch that follows the [blocks, 169] partial rows in `dpos_partial`
// [blocks][169] partial rows (padded to a 16-B multiple: the slots are updated with 16-B vector reductions), then the slots
static long long dpos_rows_floats(int blocks) { return (1LL * blocks * kBins + 3) / 4 * 4; }
extern "C" long long b200_window_attn_bwd_scratch_floats(int blocks) { return dpos_rows_floats(blocks) + 1LL * blocks * (2 * kPairs * kSlots * 32); }

extern "C" int b200_window_attn_bwd(const void* qkv, const float* pos, const float* lse, const void* dout,
                                    void* dqkv, float* dpos, float* dpos_partial, int accumulate_dpos, int B, int H, int W,
                                    int C, int heads, int shifted, void* stream) {
  int rc = check_shape(B, H, W, C, heads);
  if (rc) return rc;
  if (B == 0) return B200_OK;
  const int blocks = b200_window_attn_bwd_blocks(B, H, W, heads);
  AttnArgs a{};
  a.qkv = reinterpret_cast<const bf16*>(qkv); a.pos = pos;
  a.lse = const_cast<float*>(lse); a.dout = reinterpret_cast<const bf16*>(dout); a.dqkv = reinterpret_cast<bf16*>(dqkv);
  a.dpos_partial = dpos_partial;
  a.dslots = dpos_partial + dpos_rows_floats(blocks);  // scratch layout: [blocks][169] partial rows, then the per-lane slots
  a.B = B; a.H = H; a.W = W; a.C = C; a.heads = heads; a.shifted = shifted; a.scale = 0.17677669529663687f;  // 32^-0.5
  a.div_heads = make_fastdiv(static_cast<uint32_t>(heads)); a.div_nww = make_fastdiv(static_cast<uint32_t>(W / kWs)); a.div_nwh = make_fastdiv(static_cast<uint32_t>(H / kWs));
  a.ntasks = 1LL * B * (H / kWs) * (W / kWs) * heads; a.rev_tasks = b200_reverse_rows();
  static bool attr = false;
  if (!attr) { B200_CHECK_CUDA(cudaFuncSetAttribute(window_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem)); attr = true; }
  auto st = reinterpret_cast<cudaStream_t>(stream);
  launch_pdl(window_attn_bwd_kernel, dim3(blocks), dim3(kBwdThreads), kBwdSmem, st, a);
  B200_LAUNCH_CHECK();
  // [blocks][169] partial rows -> the 13 x 13 table gradient, fixed order (recorded, not launched, inside a reduce batch)
  return reduce_or_defer(dpos_partial, &dpos, 1, kBins, blocks, accumulate_dpos, st, kBins);
}
