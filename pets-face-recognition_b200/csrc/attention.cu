// Fused (shifted-)window attention, forward and backward, for the berniwal-variant Swin block
// (reference models/swin.py:101-135).  One warp owns one (window, head) problem: 49 tokens x 32 dims,
// padded to 64 x 32, entirely in registers + a private smem slab:
//
//   forward : S = Q K^T * scale + relpos + shift masks -> softmax -> O = P V            (P never leaves registers)
//   backward: recompute P from the saved row log-sum-exp; dV = P^T dO, dP = dO V^T,
//             dS = P o (dP - rowsum(dO o O)), dQ = scale dS K, dK = scale dS^T Q, dpos[bin] += dS
//
// The cyclic shift (roll -3 / +3) and the window partition are pure addressing here: token (r, c) of
// window (wy, wx) lives at pixel ((7 wy + r + off) mod H, (7 wx + c + off) mod W), off = 3 for shifted
// blocks, for the loads of q/k/v AND for the store of the result, so no rolled / rearranged copy of the
// activations is ever materialised.  q/k/v are read straight out of the [tokens, 3C] output of the qkv
// GEMM ([q|k|v] chunks, (head, dim) inside a chunk).
//
// 49-token problems are far below the 64/128-row tcgen05 atom and carry 3 % of the network's FLOPs, so
// the matmuls use warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate); the kernel is HBM-bound:
// it reads q, k, v once and writes o once (64-B row segments, 16 B per lane).
#include "common.cuh"

#include "b200_fe.h"

namespace {

constexpr int kWs = 7;
constexpr int kWt = 49;
constexpr int kHd = 32;
constexpr int kPitch = 40;                         // bf16 elements per smem row (80 B: conflict-free ldmatrix)
constexpr int kTileBytes = 64 * kPitch * 2;        // 5120
constexpr int kWarps = 4;
constexpr int kBins = 169;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnArgs {
  const bf16* qkv;      // [B*H*W, 3C]
  bf16* out;            // fwd: [B*H*W, C] attention output (input of to_out)
  float* lse;           // [B*H*W, heads] row log-sum-exp (nullable in inference)
  const float* pos;     // [13*13] relative position table of this block
  const bf16* dout;     // bwd: grad wrt attention output [B*H*W, C]
  const bf16* o;        // bwd: saved attention output
  bf16* dqkv;           // bwd: [B*H*W, 3C]
  float* dpos_partial;  // bwd: [gridDim.x, 169]
  int B, H, W, C, heads, shifted;
  float scale;
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
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

struct Task {
  int b, wy, wx, h;
  bool ul, lr;   // window gets the upper/lower resp. left/right shift mask (last row / last column of windows)
};

__device__ __forceinline__ Task decode_task(const AttnArgs& a, long long task) {
  const int nww = a.W / kWs, nwh = a.H / kWs;
  Task t;
  t.h = static_cast<int>(task % a.heads);
  long long win = task / a.heads;
  t.wx = static_cast<int>(win % nww);
  t.wy = static_cast<int>((win / nww) % nwh);
  t.b = static_cast<int>(win / (1LL * nww * nwh));
  t.ul = a.shifted && (t.wy == nwh - 1);
  t.lr = a.shifted && (t.wx == nww - 1);
  return t;
}

// global token row of window token i (0..48)
__device__ __forceinline__ long long token_row(const AttnArgs& a, const Task& t, int i) {
  const int off = a.shifted ? kWs / 2 : 0;
  const int r = i / kWs, c = i - r * kWs;
  int y = t.wy * kWs + r + off;
  int x = t.wx * kWs + c + off;
  if (y >= a.H) y -= a.H;
  if (x >= a.W) x -= a.W;
  return (1LL * t.b * a.H + y) * a.W + x;
}

// additive score term for (query i, key j): relpos bias, -inf on masked / padded keys; padded query rows
// get a harmless finite value.  models/swin.py:117-124.
__device__ __forceinline__ float score_bias(const float* pos_s, const Task& t, int i, int j) {
  if (j >= kWt) return -INFINITY;
  if (i >= kWt) return 0.f;
  const int ri = i / kWs, ci = i - ri * kWs;
  const int rj = j / kWs, cj = j - rj * kWs;
  constexpr int kSplit = kWs - kWs / 2;   // 4: rows/cols >= 4 are the wrapped part
  if (t.ul && ((ri >= kSplit) != (rj >= kSplit))) return -INFINITY;
  if (t.lr && ((ci >= kSplit) != (cj >= kSplit))) return -INFINITY;
  return pos_s[(rj - ri + kWs - 1) * (2 * kWs - 1) + (cj - ci + kWs - 1)];
}

// load one 49 x 32 bf16 tile (row i of the tile = window token i, 64 B = 4 lanes x 16 B) into a padded slab
__device__ __forceinline__ void load_tile_async(uint32_t slab, const bf16* base, long long ld, int col0, const long long* rows_s, int lane) {
#pragma unroll
  for (int it = 0; it < 7; ++it) {
    const int row = it * 8 + (lane >> 2);
    if (row < kWt) cp_async16(slab + row * (kPitch * 2) + (lane & 3) * 16, base + rows_s[row] * ld + col0 + (lane & 3) * 8);
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32) window_attn_fwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  float* pos_s = reinterpret_cast<float*>(smem);                                  // 169 floats (pad to 704 B)
  long long* rows_all = reinterpret_cast<long long*>(smem + 704);                 // [kWarps][64]
  uint8_t* slabs = smem + 704 + kWarps * 64 * 8;
  uint8_t* my = slabs + warp * 3 * kTileBytes;
  const uint32_t qs = smem_u32(my), ks = qs + kTileBytes, vs = ks + kTileBytes;
  long long* rows_s = rows_all + warp * 64;

  for (int i = threadIdx.x; i < kBins; i += blockDim.x) pos_s[i] = a.pos[i];
  for (int i = lane; i < 3 * kTileBytes / 16; i += 32) reinterpret_cast<uint4*>(my)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  const long long ntasks = 1LL * a.B * (a.H / kWs) * (a.W / kWs) * a.heads;
  const long long ld_qkv = 3LL * a.C;
  for (long long task = 1LL * blockIdx.x * kWarps + warp; task < ntasks; task += 1LL * gridDim.x * kWarps) {
    const Task t = decode_task(a, task);
    for (int i = lane; i < kWt; i += 32) rows_s[i] = token_row(a, t, i);
    __syncwarp();
    load_tile_async(qs, a.qkv, ld_qkv, t.h * kHd, rows_s, lane);
    load_tile_async(ks, a.qkv, ld_qkv, a.C + t.h * kHd, rows_s, lane);
    load_tile_async(vs, a.qkv, ld_qkv, 2 * a.C + t.h * kHd, rows_s, lane);
    cp_async_wait_all();
    __syncwarp();

    // K as the B operand of S = Q K^T (n = key j, k = dim d): non-transposed ldmatrix of the [j][d] slab
    uint32_t kf[8][2][2];
#pragma unroll
    for (int np = 0; np < 4; ++np)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
        ldsm_x4(ks + (np * 16 + (lane & 7) + (lane >> 4) * 8) * (kPitch * 2) + (kk * 16 + ((lane >> 3) & 1) * 8) * 2,
                kf[2 * np][kk][0], kf[2 * np][kk][1], kf[2 * np + 1][kk][0], kf[2 * np + 1][kk][1]);
    // V as the B operand of O = P V (k = key j, n = dim d): transposed ldmatrix of the [j][d] slab
    uint32_t vf[4][4][2];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int np = 0; np < 2; ++np)
        ldsm_x4_t(vs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * (kPitch * 2) + ((np * 2 + (lane >> 4)) * 8) * 2,
                  vf[kk][2 * np][0], vf[kk][2 * np][1], vf[kk][2 * np + 1][0], vf[kk][2 * np + 1][1]);

#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      float s[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(qs + (mt * 16 + (lane & 15)) * (kPitch * 2) + (kk * 16 + (lane >> 4) * 8) * 2, a0, a1, a2, a3);
#pragma unroll
        for (int n = 0; n < 8; ++n) mma16816(s[n], a0, a1, a2, a3, kf[n][kk][0], kf[n][kk][1]);
      }
      const int i0 = mt * 16 + g, i1 = i0 + 8;
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int j = n * 8 + tq * 2;
        s[n][0] = s[n][0] * a.scale + score_bias(pos_s, t, i0, j);
        s[n][1] = s[n][1] * a.scale + score_bias(pos_s, t, i0, j + 1);
        s[n][2] = s[n][2] * a.scale + score_bias(pos_s, t, i1, j);
        s[n][3] = s[n][3] * a.scale + score_bias(pos_s, t, i1, j + 1);
        m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
        m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      float l0 = 0.f, l1 = 0.f;
      uint32_t pf[8][2];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const float p0 = exp2f((s[n][0] - m0) * kLog2e), p1 = exp2f((s[n][1] - m0) * kLog2e);
        const float p2 = exp2f((s[n][2] - m1) * kLog2e), p3 = exp2f((s[n][3] - m1) * kLog2e);
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
        for (int n = 0; n < 4; ++n)
          mma16816(o[n], pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1], vf[kk][n][0], vf[kk][n][1]);
      const float r0 = 1.0f / l0, r1 = 1.0f / l1;
      __syncwarp();   // every lane has finished reading this m-tile's Q rows; reuse them as the O staging rows
      bf16* qrow = reinterpret_cast<bf16*>(my);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        if (i0 < kWt) *reinterpret_cast<uint32_t*>(qrow + i0 * kPitch + n * 8 + tq * 2) = pack_bf16(o[n][0] * r0, o[n][1] * r0);
        if (i1 < kWt) *reinterpret_cast<uint32_t*>(qrow + i1 * kPitch + n * 8 + tq * 2) = pack_bf16(o[n][2] * r1, o[n][3] * r1);
      }
      if (a.lse != nullptr && tq == 0) {
        if (i0 < kWt) a.lse[rows_s[i0] * a.heads + t.h] = m0 + __logf(l0);
        if (i1 < kWt) a.lse[rows_s[i1] * a.heads + t.h] = m1 + __logf(l1);
      }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 7; ++it) {
      const int row = it * 8 + (lane >> 2);
      if (row < kWt) {
        const uint4 v = *reinterpret_cast<const uint4*>(my + row * (kPitch * 2) + (lane & 3) * 16);
        *reinterpret_cast<uint4*>(a.out + rows_s[row] * a.C + t.h * kHd + (lane & 3) * 8) = v;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32) window_attn_bwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  float* pos_s = reinterpret_cast<float*>(smem);                                  // [169] (704 B)
  float* bins = reinterpret_cast<float*>(smem + 704);                             // [169] (704 B)
  long long* rows_all = reinterpret_cast<long long*>(smem + 1408);                // [kWarps][64]
  float* stat_all = reinterpret_cast<float*>(smem + 1408 + kWarps * 64 * 8);      // [kWarps][2][64]: lse, D
  uint8_t* slabs = smem + 1408 + kWarps * 64 * 8 + kWarps * 2 * 64 * 4;
  uint8_t* my = slabs + warp * 5 * kTileBytes;                                    // q, k, v, dO, staging
  const uint32_t qs = smem_u32(my), ks = qs + kTileBytes, vs = ks + kTileBytes, dos = vs + kTileBytes;
  bf16* stage = reinterpret_cast<bf16*>(my + 4 * kTileBytes);
  long long* rows_s = rows_all + warp * 64;
  float* lse_s = stat_all + warp * 128;
  float* dsum_s = lse_s + 64;

  for (int i = threadIdx.x; i < kBins; i += blockDim.x) { pos_s[i] = a.pos[i]; bins[i] = 0.f; }
  for (int i = lane; i < 5 * kTileBytes / 16; i += 32) reinterpret_cast<uint4*>(my)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  const long long ntasks = 1LL * a.B * (a.H / kWs) * (a.W / kWs) * a.heads;
  const long long ld_qkv = 3LL * a.C;
  for (long long task = 1LL * blockIdx.x * kWarps + warp; task < ntasks; task += 1LL * gridDim.x * kWarps) {
    const Task t = decode_task(a, task);
    for (int i = lane; i < 64; i += 32) {
      if (i < kWt) {
        const long long r = token_row(a, t, i);
        rows_s[i] = r;
        lse_s[i] = a.lse[r * a.heads + t.h];
      } else {
        lse_s[i] = 0.f;
        dsum_s[i] = 0.f;
      }
    }
    __syncwarp();
    load_tile_async(qs, a.qkv, ld_qkv, t.h * kHd, rows_s, lane);
    load_tile_async(ks, a.qkv, ld_qkv, a.C + t.h * kHd, rows_s, lane);
    load_tile_async(vs, a.qkv, ld_qkv, 2 * a.C + t.h * kHd, rows_s, lane);
    load_tile_async(dos, a.dout, a.C, t.h * kHd, rows_s, lane);
    // D_i = sum_d dO[i,d] * O[i,d]  (== rowsum(dP o P)); straight from global while the tiles stream in
#pragma unroll
    for (int it = 0; it < 7; ++it) {
      const int row = it * 8 + (lane >> 2);
      float acc = 0.f;
      if (row < kWt) {
        const long long off = rows_s[row] * a.C + t.h * kHd + (lane & 3) * 8;
        const uint4 u = *reinterpret_cast<const uint4*>(a.dout + off);
        const uint4 w = *reinterpret_cast<const uint4*>(a.o + off);
        float2 x, y;
        x = unpack_bf16(u.x); y = unpack_bf16(w.x); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16(u.y); y = unpack_bf16(w.y); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16(u.z); y = unpack_bf16(w.z); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16(u.w); y = unpack_bf16(w.w); acc += x.x * y.x + x.y * y.y;
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (row < kWt && (lane & 3) == 0) dsum_s[row] = acc;
    }
    cp_async_wait_all();
    __syncwarp();

    // ---- pass A: key-major (rows = keys j): dV = P^T dO, dK = scale * dS^T Q ---------------------------
#pragma unroll 1
    for (int jt = 0; jt < 4; ++jt) {
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) { st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f; dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t k0, k1, k2, k3, v0, v1, v2, v3;
        const uint32_t aoff = (jt * 16 + (lane & 15)) * (kPitch * 2) + (kk * 16 + (lane >> 4) * 8) * 2;
        ldsm_x4(ks + aoff, k0, k1, k2, k3);     // A = K rows (keys)
        ldsm_x4(vs + aoff, v0, v1, v2, v3);     // A = V rows (keys)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b0, b1, b2, b3;
          const uint32_t boff = (np * 16 + (lane & 7) + (lane >> 4) * 8) * (kPitch * 2) + (kk * 16 + ((lane >> 3) & 1) * 8) * 2;
          ldsm_x4(qs + boff, b0, b1, b2, b3);   // B = Q  (n = query i, k = d)
          mma16816(st[2 * np], k0, k1, k2, k3, b0, b1);
          mma16816(st[2 * np + 1], k0, k1, k2, k3, b2, b3);
          ldsm_x4(dos + boff, b0, b1, b2, b3);  // B = dO (n = query i, k = d)
          mma16816(dpt[2 * np], v0, v1, v2, v3, b0, b1);
          mma16816(dpt[2 * np + 1], v0, v1, v2, v3, b2, b3);
        }
      }
      const int j0 = jt * 16 + g, j1 = j0 + 8;
      uint32_t pf[8][2], df[8][2];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int i = n * 8 + tq * 2;
        float p[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ii = i + (e & 1), jj = (e < 2) ? j0 : j1;
          const bool valid = ii < kWt && jj < kWt;
          const float sc = st[n][e] * a.scale + score_bias(pos_s, t, ii, jj);
          p[e] = valid ? exp2f((sc - lse_s[ii]) * kLog2e) : 0.f;
          ds[e] = p[e] * (dpt[n][e] - dsum_s[ii]) * a.scale;
        }
        pf[n][0] = pack_bf16(p[0], p[1]); pf[n][1] = pack_bf16(p[2], p[3]);
        df[n][0] = pack_bf16(ds[0], ds[1]); df[n][1] = pack_bf16(ds[2], ds[3]);
      }
      float dv[4][4], dk[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) { dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f; dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)      // contraction over queries i, 16 per step
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const uint32_t boff = (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * (kPitch * 2) + ((np * 2 + (lane >> 4)) * 8) * 2;
          ldsm_x4_t(dos + boff, b0, b1, b2, b3);   // B = dO (k = i, n = d)
          mma16816(dv[2 * np], pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1], b0, b1);
          mma16816(dv[2 * np + 1], pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1], b2, b3);
          ldsm_x4_t(qs + boff, b0, b1, b2, b3);    // B = Q  (k = i, n = d)
          mma16816(dk[2 * np], df[2 * kk][0], df[2 * kk][1], df[2 * kk + 1][0], df[2 * kk + 1][1], b0, b1);
          mma16816(dk[2 * np + 1], df[2 * kk][0], df[2 * kk][1], df[2 * kk + 1][0], df[2 * kk + 1][1], b2, b3);
        }
      // stage dK rows then dV rows of this key tile and write them out (64-B row segments)
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const float(&src)[4] = which == 0 ? dk[n] : dv[n];
          *reinterpret_cast<uint32_t*>(stage + g * kPitch + n * 8 + tq * 2) = pack_bf16(src[0], src[1]);
          *reinterpret_cast<uint32_t*>(stage + (g + 8) * kPitch + n * 8 + tq * 2) = pack_bf16(src[2], src[3]);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int r = it * 8 + (lane >> 2);
          const int j = jt * 16 + r;
          if (j < kWt) {
            const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(stage) + r * (kPitch * 2) + (lane & 3) * 16);
            *reinterpret_cast<uint4*>(a.dqkv + rows_s[j] * ld_qkv + (which == 0 ? 1 : 2) * a.C + t.h * kHd + (lane & 3) * 8) = v;
          }
        }
      }
    }

    // ---- pass B: query-major (rows = queries i): dQ = scale * dS K, dpos[bin(i,j)] += dS ----------------
#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      float s[8][4], dp[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t q0, q1, q2, q3, d0, d1, d2, d3;
        const uint32_t aoff = (mt * 16 + (lane & 15)) * (kPitch * 2) + (kk * 16 + (lane >> 4) * 8) * 2;
        ldsm_x4(qs + aoff, q0, q1, q2, q3);
        ldsm_x4(dos + aoff, d0, d1, d2, d3);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b0, b1, b2, b3;
          const uint32_t boff = (np * 16 + (lane & 7) + (lane >> 4) * 8) * (kPitch * 2) + (kk * 16 + ((lane >> 3) & 1) * 8) * 2;
          ldsm_x4(ks + boff, b0, b1, b2, b3);   // B = K (n = key j, k = d)
          mma16816(s[2 * np], q0, q1, q2, q3, b0, b1);
          mma16816(s[2 * np + 1], q0, q1, q2, q3, b2, b3);
          ldsm_x4(vs + boff, b0, b1, b2, b3);   // B = V (n = key j, k = d)
          mma16816(dp[2 * np], d0, d1, d2, d3, b0, b1);
          mma16816(dp[2 * np + 1], d0, d1, d2, d3, b2, b3);
        }
      }
      const int i0 = mt * 16 + g, i1 = i0 + 8;
      uint32_t df[8][2];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int j = n * 8 + tq * 2;
        float ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ii = (e < 2) ? i0 : i1, jj = j + (e & 1);
          const bool valid = ii < kWt && jj < kWt;
          const float sc = s[n][e] * a.scale + score_bias(pos_s, t, ii, jj);
          const float p = valid ? exp2f((sc - lse_s[ii]) * kLog2e) : 0.f;
          ds[e] = p * (dp[n][e] - dsum_s[ii]);
          if (valid && p != 0.f) {
            const int ri = ii / kWs, ci = ii - ri * kWs, rj = jj / kWs, cj = jj - rj * kWs;
            atomicAdd(&bins[(rj - ri + kWs - 1) * (2 * kWs - 1) + (cj - ci + kWs - 1)], ds[e]);
          }
          ds[e] *= a.scale;
        }
        df[n][0] = pack_bf16(ds[0], ds[1]); df[n][1] = pack_bf16(ds[2], ds[3]);
      }
      float dq[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) { dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)      // contraction over keys j
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const uint32_t boff = (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * (kPitch * 2) + ((np * 2 + (lane >> 4)) * 8) * 2;
          ldsm_x4_t(ks + boff, b0, b1, b2, b3);    // B = K (k = j, n = d)
          mma16816(dq[2 * np], df[2 * kk][0], df[2 * kk][1], df[2 * kk + 1][0], df[2 * kk + 1][1], b0, b1);
          mma16816(dq[2 * np + 1], df[2 * kk][0], df[2 * kk][1], df[2 * kk + 1][0], df[2 * kk + 1][1], b2, b3);
        }
      __syncwarp();
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        *reinterpret_cast<uint32_t*>(stage + g * kPitch + n * 8 + tq * 2) = pack_bf16(dq[n][0], dq[n][1]);
        *reinterpret_cast<uint32_t*>(stage + (g + 8) * kPitch + n * 8 + tq * 2) = pack_bf16(dq[n][2], dq[n][3]);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int r = it * 8 + (lane >> 2);
        const int i = mt * 16 + r;
        if (i < kWt) {
          const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(stage) + r * (kPitch * 2) + (lane & 3) * 16);
          *reinterpret_cast<uint4*>(a.dqkv + rows_s[i] * ld_qkv + t.h * kHd + (lane & 3) * 8) = v;
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) a.dpos_partial[1LL * blockIdx.x * kBins + i] = bins[i];
}

__global__ void dpos_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int blocks, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kBins) return;
  float acc = accumulate ? out[i] : 0.f;
  for (int b = 0; b < blocks; ++b) acc += partial[1LL * b * kBins + i];
  out[i] = acc;
}

constexpr int kFwdSmem = 704 + kWarps * 64 * 8 + kWarps * 3 * kTileBytes;
constexpr int kBwdSmem = 1408 + kWarps * 64 * 8 + kWarps * 2 * 64 * 4 + kWarps * 5 * kTileBytes;

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
  static bool attr = false;
  if (!attr) { B200_CHECK_CUDA(cudaFuncSetAttribute(window_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem)); attr = true; }
  const long long ntasks = 1LL * B * (H / kWs) * (W / kWs) * heads;
  long long blocks = (ntasks + kWarps - 1) / kWarps;
  const long long cap = 1LL * b200_num_sms() * 3 * 4;
  if (blocks > cap) blocks = cap;
  window_attn_fwd_kernel<<<static_cast<unsigned>(blocks), kWarps * 32, kFwdSmem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_window_attn_bwd_blocks(int B, int H, int W, int heads) {
  const long long ntasks = 1LL * B * (H / kWs) * (W / kWs) * heads;
  long long blocks = (ntasks + kWarps - 1) / kWarps;
  const long long cap = 1LL * b200_num_sms() * 2;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : static_cast<int>(blocks);
}

extern "C" int b200_window_attn_bwd(const void* qkv, const float* pos, const void* o, const float* lse, const void* dout,
                                    void* dqkv, float* dpos, float* dpos_partial, int accumulate_dpos, int B, int H, int W,
                                    int C, int heads, int shifted, void* stream) {
  int rc = check_shape(B, H, W, C, heads);
  if (rc) return rc;
  if (B == 0) return B200_OK;
  AttnArgs a{};
  a.qkv = reinterpret_cast<const bf16*>(qkv); a.pos = pos; a.o = reinterpret_cast<const bf16*>(o);
  a.lse = const_cast<float*>(lse); a.dout = reinterpret_cast<const bf16*>(dout); a.dqkv = reinterpret_cast<bf16*>(dqkv);
  a.dpos_partial = dpos_partial;
  a.B = B; a.H = H; a.W = W; a.C = C; a.heads = heads; a.shifted = shifted; a.scale = 0.17677669529663687f;
  static bool attr = false;
  if (!attr) { B200_CHECK_CUDA(cudaFuncSetAttribute(window_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem)); attr = true; }
  const int blocks = b200_window_attn_bwd_blocks(B, H, W, heads);
  auto st = reinterpret_cast<cudaStream_t>(stream);
  window_attn_bwd_kernel<<<blocks, kWarps * 32, kBwdSmem, st>>>(a);
  B200_LAUNCH_CHECK();
  dpos_reduce_kernel<<<1, 192, 0, st>>>(dpos_partial, dpos, blocks, accumulate_dpos);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
