// HBM-bound kernels of the Swin path: LayerNorm fwd/bwd, patch gather (nn.Unfold order), mean pool,
// bf16 transpose, column sums (bias gradients), weight casts, fused multi-tensor SGD/AdamW.
// All are coalesced / 16-B vectorised, fp32 statistics, bf16 storage.
#include "common.cuh"

#include "b200_fe.h"

namespace {

__device__ __forceinline__ void ld8(const bf16* p, float* x) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = unpack_bf16(u.x); x[0] = f.x; x[1] = f.y;
  f = unpack_bf16(u.y); x[2] = f.x; x[3] = f.y;
  f = unpack_bf16(u.z); x[4] = f.x; x[5] = f.y;
  f = unpack_bf16(u.w); x[6] = f.x; x[7] = f.y;
}
__device__ __forceinline__ uint4 ld_nc_v4(const bf16* p) {     // streaming 16-B load: read once, do not keep in L1
  uint4 u;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  return u;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* x) {
  float2 f;
  f = unpack_bf16(u.x); x[0] = f.x; x[1] = f.y;
  f = unpack_bf16(u.y); x[2] = f.x; x[3] = f.y;
  f = unpack_bf16(u.z); x[4] = f.x; x[5] = f.y;
  f = unpack_bf16(u.w); x[6] = f.x; x[7] = f.y;
}
__device__ __forceinline__ void st8(bf16* p, const float* x) {
  uint4 u;
  u.x = pack_bf16(x[0], x[1]); u.y = pack_bf16(x[2], x[3]); u.z = pack_bf16(x[4], x[5]); u.w = pack_bf16(x[6], x[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void ld8f(const float* p, float* x) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm forward (nn.LayerNorm(C), eps 1e-5, models/swin.py:29,215).  LPR lanes cooperate on one
// row, each lane holding up to MAXIT chunks of 8 channels in registers; two-pass statistics in fp32.
// ---------------------------------------------------------------------------------------------
template <int LPR, int MAXIT, int U>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, bf16* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd,
                                                            long long M, int C, float eps, int rev, const WinMap wm) {
  pdl_grid_sync();
  // rev: walk the rows from the last to the first.  The producer of x (a GEMM, tiles in ascending row order) has just
  // left its last ~100 MB in L2 and the consumer of y (the next GEMM) starts at row 0: running this kernel backwards turns
  // both hand-overs into L2 hits when the tensor is larger than L2 (stages 1-2).
  auto phys = [&](long long row) { return rev ? M - 1 - row : row; };
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const long long warp_global = (1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (1LL * gridDim.x * blockDim.x) >> 5;
  const int chunks = C >> 3;
  // U row groups per iteration: all their loads are issued before any arithmetic (memory-level parallelism)
  for (long long row0 = warp_global * (RPW * U); row0 < M; row0 += nwarps * (RPW * U)) {
    uint4 raw[U][MAXIT];                 // kept packed until used: U * MAXIT 16-B loads in flight per lane
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * RPW + sub;
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        const int ch = l + it * LPR;
        raw[u][it] = (row < M && ch < chunks) ? ld_nc_v4(x + phys(row) * C + ch * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * RPW + sub;
      const bool live = row < M;
      float v[MAXIT][8];
      float s = 0.f;
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        unpack8(raw[u][it], v[it]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[it][i];
      }
      const float mu = group_sum<LPR>(s) / C;
      float q = 0.f;
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        const int ch = l + it * LPR;
        if (ch < chunks) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float d = v[it][i] - mu; q += d * d; }
        }
      }
      const float rs = rsqrtf(group_sum<LPR>(q) / C + eps);
      if (live) {
        // wm: the normalised rows leave in window-major order (the attention kernels' operand layout, see WinMap)
        const long long orow = wm.enabled ? win_row(wm, phys(row)) : phys(row);
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
          const int ch = l + it * LPR;
          if (ch < chunks) {
            float g[8], b[8], o[8];
            ld8f(gamma + ch * 8, g);
            ld8f(beta + ch * 8, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = (v[it][i] - mu) * rs * g[i] + b[i];
            st8(y + orow * C + ch * 8, o);
          }
        }
        if (l == 0) {
          if (mean) mean[phys(row)] = mu;
          if (rstd) rstd[phys(row)] = rs;
        }
      }
    }
  }
}

// LayerNorm backward.  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; optional
// fused residual-gradient add (dx_out = dres_in + dx).  dgamma/dbeta - and, with WITH_RES, the column sums of
// dres_in, which are the bias gradient of the Linear that produced the residual branch (fc2 / to_out) - are
// register partials per thread over its rows, CTA-reduced through smem, one [3C] partial row per CTA
// (reduced by splitk_reduce in a fixed order -> deterministic).
constexpr int kLnBwdThreads = 128;
// resident CTAs per SM the register budget is tuned for (two row groups of raw loads + the column partials per thread)
constexpr int ln_bwd_ctas_per_sm(int maxit) { return maxit == 1 ? 4 : (maxit == 2 ? 3 : (maxit == 3 ? 2 : 1)); }

template <int LPR, int MAXIT, bool WITH_RES>
__global__ void __launch_bounds__(kLnBwdThreads, ln_bwd_ctas_per_sm(MAXIT)) layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const bf16* dres,
                                                            bf16* dx, float* __restrict__ partial,
                                                            long long M, int C, int rev, const WinMap wm) {
  pdl_grid_sync();
  constexpr int RPW = 32 / LPR;
  extern __shared__ float red[];   // [warps][3][C]
  auto phys = [&](long long row) { return rev ? M - 1 - row : row; };     // see layernorm_fwd_kernel
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR;
  const long long warp_global = (1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (1LL * gridDim.x * blockDim.x) >> 5;
  const int chunks = C >> 3;
  float dg[MAXIT][8], db[MAXIT][8], dr[WITH_RES ? MAXIT : 1][8];
#pragma unroll
  for (int it = 0; it < MAXIT; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) { dg[it][i] = 0.f; db[it][i] = 0.f; if (WITH_RES) dr[it][i] = 0.f; }

  // software pipeline: the raw (packed bf16) loads of the warp's next D row groups are in flight while the current one is
  // reduced (D slots, each refilled as soon as it has been consumed), so every lane keeps D * MAXIT * 3 16-B loads
  // outstanding - the narrow layers (C = 96 / 192: one chunk per lane) need the depth to cover the HBM latency
  constexpr int D = MAXIT == 1 ? 4 : 1;
  uint4 rd[D][MAXIT], rx[D][MAXIT], rr[D][MAXIT];
  float mu_r[D], rs_r[D];
  auto fetch = [&](int slot, long long row) {
    const bool live = row < M;
    const long long prow = phys(row);
    mu_r[slot] = live ? mean[prow] : 0.f;
    rs_r[slot] = live ? rstd[prow] : 0.f;
    const long long drow = (live && wm.enabled) ? win_row(wm, prow) : prow;      // wm: dy arrives in window-major order
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int ch = l + it * LPR;
      if (live && ch < chunks) {
        rd[slot][it] = ld_nc_v4(dy + drow * C + ch * 8);
        rx[slot][it] = ld_nc_v4(x + prow * C + ch * 8);
        if (dres) rr[slot][it] = ld_nc_v4(dres + prow * C + ch * 8);
      }
    }
  };
  const long long stride = nwarps * RPW;
  long long row0 = warp_global * RPW;
#pragma unroll
  for (int d = 0; d < D; ++d) fetch(d, row0 + d * stride + sub);
  for (; row0 < M; row0 += D * stride) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const long long row = row0 + d * stride + sub;
      const bool live = row < M;
      const float mu = mu_r[d], rs = rs_r[d];
      uint4 cd[MAXIT], cx[MAXIT], cr[MAXIT];
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) { cd[it] = rd[d][it]; cx[it] = rx[d][it]; cr[it] = rr[d][it]; }
      fetch(d, row0 + (d + D) * stride + sub);
      float g[MAXIT][8], xh[MAXIT][8];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        const int ch = l + it * LPR;
        if (live && ch < chunks) {
          float dv[8], xv[8], gm[8];
          unpack8(cd[it], dv);
          unpack8(cx[it], xv);
          ld8f(gamma + ch * 8, gm);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            xh[it][i] = (xv[i] - mu) * rs;
            g[it][i] = dv[i] * gm[i];
            s1 += g[it][i];
            s2 += g[it][i] * xh[it][i];
            dg[it][i] += dv[i] * xh[it][i];
            db[it][i] += dv[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) { g[it][i] = 0.f; xh[it][i] = 0.f; }
        }
      }
      s1 = group_sum<LPR>(s1) / C;
      s2 = group_sum<LPR>(s2) / C;
      if (live) {
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
          const int ch = l + it * LPR;
          if (ch < chunks) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = rs * (g[it][i] - s1 - xh[it][i] * s2);
            if (dres) {
              float r[8];
              unpack8(cr[it], r);
#pragma unroll
              for (int i = 0; i < 8; ++i) { o[i] += r[i]; if (WITH_RES) dr[it][i] += r[i]; }
            }
            st8(dx + phys(row) * C + ch * 8, o);
          }
        }
      }
    }
  }
  // fold the RPW sub-rows of a warp, then the warps of the CTA
#pragma unroll
  for (int it = 0; it < MAXIT; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int o = 16; o >= LPR; o >>= 1) {
        dg[it][i] += __shfl_xor_sync(0xffffffffu, dg[it][i], o);
        db[it][i] += __shfl_xor_sync(0xffffffffu, db[it][i], o);
        if (WITH_RES) dr[it][i] += __shfl_xor_sync(0xffffffffu, dr[it][i], o);
      }
    }
  const int warps = blockDim.x >> 5;
  if (sub == 0) {
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int ch = l + it * LPR;
      if (ch < chunks) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          red[(warp * 3 + 0) * C + ch * 8 + i] = dg[it][i];
          red[(warp * 3 + 1) * C + ch * 8 + i] = db[it][i];
          red[(warp * 3 + 2) * C + ch * 8 + i] = WITH_RES ? dr[it][i] : 0.f;
        }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * C; c += blockDim.x) {
    const int which = c / C, col = c % C;
    float a = 0.f;
    for (int w = 0; w < warps; ++w) a += red[(w * 3 + which) * C + col];
    partial[1LL * blockIdx.x * 3 * C + c] = a;
  }
}

// ---------------------------------------------------------------------------------------------
// Patch gather == nn.Unfold(k = s = df) + NHWC view (models/swin.py:162-166): out[m, c*df*df + kh*df + kw]
// ---------------------------------------------------------------------------------------------
// stage 1: fp32 NCHW image, df = 4 -> one thread per (output row, c, kh): float4 in, 4 bf16 out
__global__ void patch_gather_image_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int Cin, int H, int W,
                                          int df, int ldo) {
  pdl_grid_sync();
  const int Ho = H / df, Wo = W / df;
  const int per_row = Cin * df;                       // (c, kh) pairs, each df(=4) contiguous kw
  const long long total = 1LL * B * Ho * Wo * per_row;
  for (long long idx = 1LL * blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += 1LL * gridDim.x * blockDim.x) {
    const int j = static_cast<int>(idx % per_row);
    const long long m = idx / per_row;
    const int c = j / df, kh = j % df;
    const int px = static_cast<int>(m % Wo);
    const int py = static_cast<int>((m / Wo) % Ho);
    const int b = static_cast<int>(m / (1LL * Wo * Ho));
    const float4 v = __ldg(reinterpret_cast<const float4*>(img + ((1LL * b * Cin + c) * H + py * df + kh) * W + px * df));
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(out + m * ldo + j * 4) = o;
  }
}

// same gather straight from uint8 pixels: the /255 of torchvision's ToTensor (configs/dog_fe/fe_dogs_config.py:25,31) is fused
// here, so a host batch crosses PCIe at 1 byte per sample instead of 4
__global__ void patch_gather_image_u8_kernel(const uint8_t* __restrict__ img, bf16* __restrict__ out, int B, int Cin, int H, int W,
                                             int df, int ldo) {
  pdl_grid_sync();
  const int Ho = H / df, Wo = W / df;
  const int per_row = Cin * df;
  const long long total = 1LL * B * Ho * Wo * per_row;
  for (long long idx = 1LL * blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += 1LL * gridDim.x * blockDim.x) {
    const int j = static_cast<int>(idx % per_row);
    const long long m = idx / per_row;
    const int c = j / df, kh = j % df;
    const int px = static_cast<int>(m % Wo);
    const int py = static_cast<int>((m / Wo) % Ho);
    const int b = static_cast<int>(m / (1LL * Wo * Ho));
    const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(img + ((1LL * b * Cin + c) * H + py * df + kh) * W + px * df));
    uint2 o;
    o.x = pack_bf16(v.x / 255.0f, v.y / 255.0f);
    o.y = pack_bf16(v.z / 255.0f, v.w / 255.0f);
    *reinterpret_cast<uint2*>(out + m * ldo + j * 4) = o;
  }
}

// stages 2-4: bf16 NHWC, df = 2 -> one thread per (output row, channel pair): 4 x bf162 in, 16 B out
// backward (scatter == exact inverse, every input pixel appears once) uses the same indexing.
template <bool BACKWARD>
__global__ void patch_gather_nhwc_kernel(bf16* __restrict__ x, bf16* __restrict__ cols, int B, int H, int W, int C) {
  pdl_grid_sync();
  const int Ho = H / 2, Wo = W / 2, cp = C / 2;
  const long long total = 1LL * B * Ho * Wo * cp;
  for (long long idx = 1LL * blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += 1LL * gridDim.x * blockDim.x) {
    const int c2 = static_cast<int>(idx % cp);
    const long long m = idx / cp;
    const int px = static_cast<int>(m % Wo);
    const int py = static_cast<int>((m / Wo) % Ho);
    const int b = static_cast<int>(m / (1LL * Wo * Ho));
    bf16* base = x + ((1LL * b * H + py * 2) * W + px * 2) * C + c2 * 2;
    bf16* dst = cols + m * (4LL * C) + c2 * 8;        // features (c, kh, kw) for c = 2*c2, 2*c2+1
    if (!BACKWARD) {
      const bf162 p00 = *reinterpret_cast<const bf162*>(base);
      const bf162 p01 = *reinterpret_cast<const bf162*>(base + C);
      const bf162 p10 = *reinterpret_cast<const bf162*>(base + 1LL * W * C);
      const bf162 p11 = *reinterpret_cast<const bf162*>(base + 1LL * W * C + C);
      bf162 o[4];
      o[0] = bf162(p00.x, p01.x); o[1] = bf162(p10.x, p11.x);   // channel 2*c2  : (kh,kw) = 00,01,10,11
      o[2] = bf162(p00.y, p01.y); o[3] = bf162(p10.y, p11.y);   // channel 2*c2+1
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(o);
    } else {
      bf162 o[4];
      *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(dst);
      *reinterpret_cast<bf162*>(base) = bf162(o[0].x, o[2].x);
      *reinterpret_cast<bf162*>(base + C) = bf162(o[0].y, o[2].y);
      *reinterpret_cast<bf162*>(base + 1LL * W * C) = bf162(o[1].x, o[3].x);
      *reinterpret_cast<bf162*>(base + 1LL * W * C + C) = bf162(o[1].y, o[3].y);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// spatial mean (models/swin.py:224) and its backward (broadcast / T)
// ---------------------------------------------------------------------------------------------
__global__ void mean_pool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int T, int C) {
  pdl_grid_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * (C / 2)) return;
  const int c2 = idx % (C / 2), b = idx / (C / 2);
  float a0 = 0.f, a1 = 0.f;
  for (int t = 0; t < T; ++t) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const bf162*>(x + (1LL * b * T + t) * C + c2 * 2));
    a0 += f.x; a1 += f.y;
  }
  *reinterpret_cast<bf162*>(y + 1LL * b * C + c2 * 2) = __floats2bfloat162_rn(a0 / T, a1 / T);
}
__global__ void mean_pool_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int B, int T, int C) {
  pdl_grid_sync();
  const long long idx = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 1LL * B * T * (C / 2)) return;
  const int c2 = static_cast<int>(idx % (C / 2));
  const int b = static_cast<int>(idx / (1LL * T * (C / 2)));
  const float2 f = __bfloat1622float2(*reinterpret_cast<const bf162*>(dy + 1LL * b * C + c2 * 2));
  *reinterpret_cast<bf162*>(dx + idx * 2) = __floats2bfloat162_rn(f.x / T, f.y / T);
}

// ---------------------------------------------------------------------------------------------
// [R, Cc] -> [Cc, R] 16-bit transpose, 64x64 tiles through padded smem (coalesced both ways)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                                                          long long R, int Cc, long long ld_in, long long ld_out) {
  pdl_grid_sync();
  __shared__ uint16_t tile[64][66];
  const long long r0 = 64LL * blockIdx.x;
  const int c0 = 64 * blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 64; i += 8) {
    const long long r = r0 + i;
    const int c = c0 + tx * 2;
    uint32_t v = 0;
    if (r < R && c < Cc) v = *reinterpret_cast<const uint32_t*>(in + r * ld_in + c);   // Cc even
    tile[i][tx * 2] = static_cast<uint16_t>(v & 0xffff);
    tile[i][tx * 2 + 1] = static_cast<uint16_t>(v >> 16);
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i;
    const long long r = r0 + tx * 2;
    if (c < Cc && r < R) {
      const uint32_t v = static_cast<uint32_t>(tile[tx * 2][i]) | (static_cast<uint32_t>(tile[tx * 2 + 1][i]) << 16);
      if (r + 1 < R) *reinterpret_cast<uint32_t*>(out + c * ld_out + r) = v;
      else out[c * ld_out + r] = static_cast<uint16_t>(v & 0xffff);
    }
  }
}

// fp32 [R, Cc] -> bf16 [R, Cc] (dst) and/or bf16 [Cc, R] (dst_t).  Used for weights (small).
__global__ void __launch_bounds__(256) cast_transpose_kernel(const float* __restrict__ in, bf16* __restrict__ dst,
                                                             bf16* __restrict__ dst_t, int R, int Cc) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int r0 = 32 * blockIdx.x, c0 = 32 * blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < Cc) {
      v = in[1LL * r * Cc + c];
      if (dst) dst[1LL * r * Cc + c] = __float2bfloat16_rn(v);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  if (dst_t) {
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;
      if (c < Cc && r < R) dst_t[1LL * c * R + r] = __float2bfloat16_rn(tile[tx][i]);
    }
  }
}

// the same for a whole table of matrices in one launch (the per-step fp32 -> bf16 weight refresh): block -> (job, tile)
__global__ void __launch_bounds__(256) cast_transpose_multi_kernel(const float* __restrict__ in_base, bf16* __restrict__ out_base,
                                                                   const CastJobs jobs) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  int j = 0;
  while (j + 1 < jobs.n && static_cast<int>(blockIdx.x) >= jobs.job[j + 1].tile0) ++j;
  const CastJob job = jobs.job[j];
  const int t = blockIdx.x - job.tile0;
  const int r0 = 32 * (t / job.tiles_c), c0 = 32 * (t % job.tiles_c);
  const float* in = in_base + job.in_off;
  bf16* dst = out_base + job.dst_off;
  bf16* dst_t = job.dst_t_off >= 0 ? out_base + job.dst_t_off : nullptr;
  const int R = job.R, Cc = job.Cc;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < Cc) {
      v = in[1LL * r * Cc + c];
      dst[1LL * r * Cc + c] = __float2bfloat16_rn(v);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  if (dst_t) {
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;
      if (c < Cc && r < R) dst_t[1LL * c * R + r] = __float2bfloat16_rn(tile[tx][i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// column sums of a bf16 [M, N] matrix (bias gradients): 256 threads = TX column lanes (16 B = 8 columns each) x
// 256/TX row lanes; each CTA reduces a slab of rows -> partial[blk][N] (fixed order: deterministic)
// ---------------------------------------------------------------------------------------------
template <int TX>
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, float* __restrict__ partial, long long M, int N,
                                                     long long ld, long long rows_per_block) {
  pdl_grid_sync();
  constexpr int TY = 256 / TX;
  __shared__ float red[TY][TX * 8 + 1];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int col = (blockIdx.y * TX + tx) * 8;
  const long long r0 = rows_per_block * blockIdx.x;
  const long long r1 = min(M, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
#pragma unroll 4
    for (long long r = r0 + ty; r < r1; r += TY) {
      float v[8];
      ld8(x + r * ld + col, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < TX * 8; c += 256) {
    const int gc = blockIdx.y * TX * 8 + c;
    if (gc < N) {
      float a = 0.f;
#pragma unroll
      for (int y = 0; y < TY; ++y) a += red[y][c];
      partial[1LL * blockIdx.x * N + gc] = a;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused multi-tensor optimizer steps.  One launch updates every parameter: a device table lists the
// tensors, a chunk table maps CTAs onto (tensor, offset) slabs of kChunk elements.
// ---------------------------------------------------------------------------------------------
constexpr int kChunk = 4096;

// Where the gradient of a step lives.  n_src == 1: in B200OptTensor.grad.  n_src > 1 (data parallel, csrc/peer.cu): every
// rank's copy sits in its own slot of the peer arena, `stride` elements apart, `shift` elements past the table's pointer (the
// arena is double-buffered); the kernel adds the slots in slot order - identical on every rank - so the gradient all-reduce
// is folded into the optimizer pass.
struct GradSrc {
  int n_src;
  long long stride, shift;
};
__device__ __forceinline__ float load_grad(const float* __restrict__ g, long long i, const GradSrc& gs) {
  const float* q = g + gs.shift + i;
  float v = q[0];
  if (gs.n_src == 8) {          // a full node: all eight loads in flight before the first add (same slot order as the loop)
    float s[7];
#pragma unroll
    for (int r = 0; r < 7; ++r) s[r] = q[(r + 1) * gs.stride];
#pragma unroll
    for (int r = 0; r < 7; ++r) v += s[r];
    return v;
  }
  for (int r = 1; r < gs.n_src; ++r) v += q[r * gs.stride];
  return v;
}

// torch.optim.SGD(momentum, dampening 0, no nesterov): g += wd * p; buf = first ? g : mom * buf + g; p -= lr * buf
__global__ void __launch_bounds__(256) sgd_kernel(const B200OptTensor* __restrict__ tensors, const int2* __restrict__ chunks,
                                                  float grad_scale, int step_offset, const GradSrc gs) {
  pdl_grid_sync();
  const int2 ck = chunks[blockIdx.x];
  const B200OptTensor t = tensors[ck.x];
  float* p = reinterpret_cast<float*>(t.param);
  const float* g = reinterpret_cast<const float*>(t.grad);
  float* buf = reinterpret_cast<float*>(t.state1);
  bf16* p16 = reinterpret_cast<bf16*>(t.param_bf16);
  const long long end = min(t.numel, 1LL * ck.y + kChunk);
  for (long long i = ck.y + threadIdx.x; i < end; i += blockDim.x) {
    float pv = p[i];
    float gv = load_grad(g, i, gs) * grad_scale + t.weight_decay * pv;
    const float bv = (t.step + step_offset) == 0 ? gv : t.beta1 * buf[i] + gv;
    buf[i] = bv;
    pv -= t.lr * bv;
    p[i] = pv;
    if (p16) p16[i] = __float2bfloat16_rn(pv);
  }
}

// torch.optim.AdamW: p *= 1 - lr*wd; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256) adamw_kernel(const B200OptTensor* __restrict__ tensors, const int2* __restrict__ chunks,
                                                    float grad_scale, int step_offset, const GradSrc gs) {
  pdl_grid_sync();
  const int2 ck = chunks[blockIdx.x];
  const B200OptTensor t = tensors[ck.x];
  float* p = reinterpret_cast<float*>(t.param);
  const float* g = reinterpret_cast<const float*>(t.grad);
  float* m = reinterpret_cast<float*>(t.state1);
  float* v = reinterpret_cast<float*>(t.state2);
  bf16* p16 = reinterpret_cast<bf16*>(t.param_bf16);
  const float stepf = static_cast<float>(t.step + step_offset + 1);
  const float bc1 = 1.0f - powf(t.beta1, stepf);
  const float bc2s = sqrtf(1.0f - powf(t.beta2, stepf));
  const long long end = min(t.numel, 1LL * ck.y + kChunk);
  for (long long i = ck.y + threadIdx.x; i < end; i += blockDim.x) {
    float pv = p[i] * (1.0f - t.lr * t.weight_decay);
    const float gv = load_grad(g, i, gs) * grad_scale;
    const float mv = t.beta1 * m[i] + (1.0f - t.beta1) * gv;
    const float vv = t.beta2 * v[i] + (1.0f - t.beta2) * gv * gv;
    m[i] = mv; v[i] = vv;
    pv -= (t.lr / bc1) * mv / (sqrtf(vv) / bc2s + t.eps);
    p[i] = pv;
    if (p16) p16[i] = __float2bfloat16_rn(pv);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  pdl_grid_sync();
  const long long i = (1LL * blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    uint2 o;
    o.x = pack_bf16(v.x, v.y); o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(out + i) = o;
  } else {
    for (long long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
  }
}

int grid_for(long long work_items, int threads, int max_blocks) {
  long long b = (work_items + threads - 1) / threads;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

template <int LPR, int MAXIT, int U = (MAXIT == 1 ? 4 : (MAXIT <= 3 ? 2 : 1))>
int ln_fwd_launch(const bf16* x, const float* g, const float* b, bf16* y, float* mean, float* rstd, long long M, int C, float eps,
                  cudaStream_t st, const WinMap& wm) {
  const int rpw = (32 / LPR) * U;
  const long long warps = (M + rpw - 1) / rpw;
  const int blocks = grid_for(warps * 32, 256, b200_num_sms() * 8);
  const bool prof = b200_prof_kind_begin(st, B200_PROF_LN_FWD, 0.0, 1.0 * M * C * 2.0 * 2.0 + 1.0 * M * 8.0);
  launch_pdl(layernorm_fwd_kernel<LPR, MAXIT, U>, dim3(blocks), dim3(256), 0, st, x, g, b, y, mean, rstd, M, C, eps, b200_reverse_rows(), wm);
  if (prof) b200_prof_kind_end(st);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
template <int LPR, int MAXIT>
int ln_bwd_launch(const bf16* dy, const bf16* x, const float* g, const float* mean, const float* rstd, const bf16* dres, bf16* dx,
                  float* partial, long long M, int C, int blocks, bool with_res, cudaStream_t st, const WinMap& wm) {
  const size_t smem = sizeof(float) * (kLnBwdThreads / 32) * 3 * C;
  const bool prof = b200_prof_kind_begin(st, B200_PROF_LN_BWD, 0.0, 1.0 * M * C * 2.0 * (dres ? 4.0 : 3.0) + 1.0 * M * 8.0);
  if (with_res) {
    if (smem > 48 * 1024)
      B200_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<LPR, MAXIT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(layernorm_bwd_kernel<LPR, MAXIT, true>, dim3(blocks), dim3(kLnBwdThreads), smem, st, dy, x, g, mean, rstd, dres, dx, partial, M, C, b200_reverse_rows(), wm);
  } else {
    if (smem > 48 * 1024)
      B200_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<LPR, MAXIT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(layernorm_bwd_kernel<LPR, MAXIT, false>, dim3(blocks), dim3(kLnBwdThreads), smem, st, dy, x, g, mean, rstd, dres, dx, partial, M, C, b200_reverse_rows(), wm);
  }
  if (prof) b200_prof_kind_end(st);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
static int layernorm_fwd_impl(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                              long long M, int C, float eps, const WinMap& wm, void* stream) {
  B200_REQUIRE(C % 8 == 0 && C >= 8 && C <= 1536, "layernorm: C=%d unsupported (multiple of 8, <= 1536)", C);
  if (M == 0) return B200_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  auto X = reinterpret_cast<const bf16*>(x);
  auto Y = reinterpret_cast<bf16*>(y);
  // widths of the form 24 * 2^k (Swin: 96 / 192 / 384 / 768): C / 24 lanes per row hold exactly three 16-B chunks each -
  // no idle lanes, short shuffle trees, and 12 loads in flight per lane
  if (C == 96) return ln_fwd_launch<4, 3, 4>(X, gamma, beta, Y, mean, rstd, M, C, eps, st, wm);
  if (C == 192) return ln_fwd_launch<8, 3, 4>(X, gamma, beta, Y, mean, rstd, M, C, eps, st, wm);
  if (C == 384) return ln_fwd_launch<16, 3, 4>(X, gamma, beta, Y, mean, rstd, M, C, eps, st, wm);
  if (C <= 128) return ln_fwd_launch<16, 1>(X, gamma, beta, Y, mean, rstd, M, C, eps, st, wm);
  if (C <= 256) return ln_fwd_launch<32, 1>(X, gamma, beta, Y, mean, rstd, M, C, eps, st, wm);
  if (C <= 512) return ln_fwd_launch<32, 2>(X, gamma, beta, Y, mean, rstd, M, C, eps, st, wm);
  if (C <= 768) return ln_fwd_launch<32, 3>(X, gamma, beta, Y, mean, rstd, M, C, eps, st, wm);
  return ln_fwd_launch<32, 6>(X, gamma, beta, Y, mean, rstd, M, C, eps, st, wm);
}

extern "C" int b200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                  long long M, int C, float eps, void* stream) {
  return layernorm_fwd_impl(x, gamma, beta, y, mean, rstd, M, C, eps, no_winmap(), stream);
}

static int check_window_grid(int B, int H, int W) {
  B200_REQUIRE(B >= 0 && H > 0 && W > 0 && H % 7 == 0 && W % 7 == 0 && 1LL * B * H * W < (1LL << 31),
               "window order: H=%d W=%d must be multiples of 7 and B*H*W < 2^31", H, W);
  return B200_OK;
}

// y rows are written in WINDOW-MAJOR order (x, mean, rstd stay in raster order)
extern "C" int b200_layernorm_fwd_windows(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                          int B, int H, int W, int C, int shifted, float eps, void* stream) {
  int rc = check_window_grid(B, H, W);
  if (rc) return rc;
  return layernorm_fwd_impl(x, gamma, beta, y, mean, rstd, 1LL * B * H * W, C, eps, make_winmap(H, W, shifted), stream);
}

// 128-thread CTAs; each CTA ends with a cross-warp reduction and one [3C] partial row, so CTAs are kept fat (>= 64
// rows) and their number at the resident capacity (register-limited, see ln_bwd_ctas_per_sm)
// widths of the form 24 * 2^k (Swin: 96 / 192 / 384): C / 24 lanes per row with three 16-B chunks each - no idle lanes and a
// quarter of the per-row bookkeeping of the one-chunk-per-lane layout (B200_LN_BWD3=0: the older layout, for A/B runs)
static bool ln_bwd_three_chunks(int C) {
  static const bool on = [] { const char* e = getenv("B200_LN_BWD3"); return e == nullptr || e[0] != '0'; }();
  return on && (C == 96 || C == 192 || C == 384);
}

extern "C" int b200_layernorm_bwd_blocks(long long M, int C) {
  long long b = (M + 63) / 64;
  const int maxit = ln_bwd_three_chunks(C) ? 3 : (C <= 256 ? 1 : (C <= 512 ? 2 : (C <= 768 ? 3 : 6)));
  const int cap = b200_num_sms() * ln_bwd_ctas_per_sm(maxit);
  if (b > cap) b = cap;
  return b < 1 ? 1 : static_cast<int>(b);
}

// partial: fp32 scratch of b200_layernorm_bwd_blocks(M, C) x 3C.  dres_colsum (optional, needs dres_in) receives the column
// sums of dres_in: the bias gradient of the Linear whose output was added to the residual stream.
static int layernorm_bwd_impl(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                              const void* dres_in, void* dx_out, float* dgamma, float* dbeta, float* dres_colsum, float* partial,
                              long long M, int C, int accumulate, const WinMap& wm, void* stream) {
  B200_REQUIRE(C % 8 == 0 && C >= 8 && C <= 1536, "layernorm_bwd: C=%d unsupported", C);
  B200_REQUIRE(dres_colsum == nullptr || dres_in != nullptr, "layernorm_bwd: dres_colsum needs dres_in");
  if (M == 0) return B200_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = b200_layernorm_bwd_blocks(M, C);
  auto DY = reinterpret_cast<const bf16*>(dy);
  auto X = reinterpret_cast<const bf16*>(x);
  auto DR = reinterpret_cast<const bf16*>(dres_in);
  auto DX = reinterpret_cast<bf16*>(dx_out);
  const bool wr = dres_colsum != nullptr;
  int rc;
  if (ln_bwd_three_chunks(C)) {
    if (C == 96) rc = ln_bwd_launch<4, 3>(DY, X, gamma, mean, rstd, DR, DX, partial, M, C, blocks, wr, st, wm);
    else if (C == 192) rc = ln_bwd_launch<8, 3>(DY, X, gamma, mean, rstd, DR, DX, partial, M, C, blocks, wr, st, wm);
    else rc = ln_bwd_launch<16, 3>(DY, X, gamma, mean, rstd, DR, DX, partial, M, C, blocks, wr, st, wm);
  }
  else if (C <= 128) rc = ln_bwd_launch<16, 1>(DY, X, gamma, mean, rstd, DR, DX, partial, M, C, blocks, wr, st, wm);
  else if (C <= 256) rc = ln_bwd_launch<32, 1>(DY, X, gamma, mean, rstd, DR, DX, partial, M, C, blocks, wr, st, wm);
  else if (C <= 512) rc = ln_bwd_launch<32, 2>(DY, X, gamma, mean, rstd, DR, DX, partial, M, C, blocks, wr, st, wm);
  else if (C <= 768) rc = ln_bwd_launch<32, 3>(DY, X, gamma, mean, rstd, DR, DX, partial, M, C, blocks, wr, st, wm);
  else rc = ln_bwd_launch<32, 6>(DY, X, gamma, mean, rstd, DR, DX, partial, M, C, blocks, wr, st, wm);
  if (rc) return rc;
  float* outs[3] = {dgamma, dbeta, dres_colsum};
  return splitk_reduce_multi(partial, outs, wr ? 3 : 2, C, blocks, accumulate, st, 3LL * C);
}

// partial: fp32 scratch of b200_layernorm_bwd_blocks(M, C) x 3C.  dres_colsum (optional, needs dres_in) receives the column
// sums of dres_in: the bias gradient of the Linear whose output was added to the residual stream.
extern "C" int b200_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                  const void* dres_in, void* dx_out, float* dgamma, float* dbeta, float* dres_colsum, float* partial,
                                  long long M, int C, int accumulate, void* stream) {
  return layernorm_bwd_impl(dy, x, gamma, mean, rstd, dres_in, dx_out, dgamma, dbeta, dres_colsum, partial, M, C, accumulate, no_winmap(), stream);
}

// dy rows are READ in window-major order (everything else in raster order): the backward of b200_layernorm_fwd_windows
extern "C" int b200_layernorm_bwd_windows(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                          const void* dres_in, void* dx_out, float* dgamma, float* dbeta, float* dres_colsum,
                                          float* partial, int B, int H, int W, int C, int shifted, int accumulate, void* stream) {
  int rc = check_window_grid(B, H, W);
  if (rc) return rc;
  return layernorm_bwd_impl(dy, x, gamma, mean, rstd, dres_in, dx_out, dgamma, dbeta, dres_colsum, partial, 1LL * B * H * W, C, accumulate,
                            make_winmap(H, W, shifted), stream);
}

// out rows = in rows permuted between raster and window-major order (row = n_words 32-bit words): to_window != 0 writes
// out[window_major(r)] = in[r], otherwise out[r] = in[window_major(r)].  A layout helper for callers that hold raster-order
// tensors (tests, stand-alone use of the attention entry points); the Swin plan never needs it.
__global__ void __launch_bounds__(256) window_rows_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long M, int n_words,
                                                          int to_window, const WinMap wm) {
  pdl_grid_sync();
  const long long total = M * n_words;
  for (long long idx = 1LL * blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += 1LL * gridDim.x * blockDim.x) {
    const long long row = idx / n_words;
    const int wd = static_cast<int>(idx - row * n_words);
    const long long w = win_row(wm, row);
    if (to_window) out[w * n_words + wd] = in[idx];
    else out[idx] = in[w * n_words + wd];
  }
}

extern "C" int b200_window_rows(const void* in, void* out, int B, int H, int W, int row_bytes, int shifted, int to_window, void* stream) {
  int rc = check_window_grid(B, H, W);
  if (rc) return rc;
  B200_REQUIRE(row_bytes > 0 && row_bytes % 4 == 0, "window_rows: row_bytes must be a positive multiple of 4");
  const long long M = 1LL * B * H * W;
  if (M == 0) return B200_OK;
  launch_pdl(window_rows_kernel, dim3(grid_for(M * (row_bytes / 4), 256, b200_num_sms() * 16)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
             reinterpret_cast<const uint32_t*>(in), reinterpret_cast<uint32_t*>(out), M, row_bytes / 4, to_window, make_winmap(H, W, shifted));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_patch_gather_image(const float* img, void* out, int B, int Cin, int H, int W, int df, long long ldo,
                                       void* stream) {
  B200_REQUIRE(df == 4 && H % 4 == 0 && W % 4 == 0, "patch_gather_image: downscaling factor 4 only (got %d)", df);
  B200_REQUIRE(ldo % 4 == 0 && ldo >= Cin * 16, "patch_gather_image: bad ldo");
  const long long total = 1LL * B * (H / 4) * (W / 4) * Cin * 4;
  if (total == 0) return B200_OK;
  launch_pdl(patch_gather_image_kernel, dim3(grid_for(total, 256, b200_num_sms() * 16)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      img, reinterpret_cast<bf16*>(out), B, Cin, H, W, df, static_cast<int>(ldo));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_patch_gather_image_u8(const unsigned char* img, void* out, int B, int Cin, int H, int W, int df, long long ldo,
                                          void* stream) {
  B200_REQUIRE(df == 4 && H % 4 == 0 && W % 4 == 0, "patch_gather_image_u8: downscaling factor 4 only (got %d)", df);
  B200_REQUIRE(ldo % 4 == 0 && ldo >= Cin * 16, "patch_gather_image_u8: bad ldo");
  const long long total = 1LL * B * (H / 4) * (W / 4) * Cin * 4;
  if (total == 0) return B200_OK;
  launch_pdl(patch_gather_image_u8_kernel, dim3(grid_for(total, 256, b200_num_sms() * 16)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
             img, reinterpret_cast<bf16*>(out), B, Cin, H, W, df, static_cast<int>(ldo));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_patch_gather_nhwc(void* x, void* cols, int B, int H, int W, int C, int backward, void* stream) {
  B200_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 2 == 0, "patch_gather_nhwc: H, W, C must be even");
  const long long total = 1LL * B * (H / 2) * (W / 2) * (C / 2);
  if (total == 0) return B200_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = grid_for(total, 256, b200_num_sms() * 16);
  if (backward) launch_pdl(patch_gather_nhwc_kernel<true>, dim3(blocks), dim3(256), 0, st, reinterpret_cast<bf16*>(x), reinterpret_cast<bf16*>(cols), B, H, W, C);
  else launch_pdl(patch_gather_nhwc_kernel<false>, dim3(blocks), dim3(256), 0, st, reinterpret_cast<bf16*>(x), reinterpret_cast<bf16*>(cols), B, H, W, C);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_mean_pool(const void* in, void* out, int B, int T, int C, int backward, void* stream) {
  B200_REQUIRE(C % 2 == 0, "mean_pool: C must be even");
  if (B == 0) return B200_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (!backward) {
    const int n = B * (C / 2);
    launch_pdl(mean_pool_fwd_kernel, dim3((n + 255) / 256), dim3(256), 0, st, reinterpret_cast<const bf16*>(in), reinterpret_cast<bf16*>(out), B, T, C);
  } else {
    const long long n = 1LL * B * T * (C / 2);
    launch_pdl(mean_pool_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, reinterpret_cast<const bf16*>(in), reinterpret_cast<bf16*>(out), B, T, C);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_transpose16(const void* in, void* out, long long R, int Cc, long long ld_in, long long ld_out, void* stream) {
  B200_REQUIRE(Cc % 2 == 0 && ld_in % 2 == 0 && ld_out % 2 == 0, "transpose16: even column count / pitches required");
  if (R == 0 || Cc == 0) return B200_OK;
  dim3 grid(static_cast<unsigned>((R + 63) / 64), (Cc + 63) / 64);
  launch_pdl(transpose16_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const uint16_t*>(in),
                                                                              reinterpret_cast<uint16_t*>(out), R, Cc, ld_in, ld_out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_cast_transpose(const float* in, void* dst, void* dst_t, int R, int Cc, void* stream) {
  if (R == 0 || Cc == 0) return B200_OK;
  dim3 grid((R + 31) / 32, (Cc + 31) / 32);
  launch_pdl(cast_transpose_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), in, reinterpret_cast<bf16*>(dst),
                                                                                 reinterpret_cast<bf16*>(dst_t), R, Cc);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int cast_transpose_multi(const float* in_base, void* out_base, const CastJobs& jobs, int total_tiles, cudaStream_t stream) {
  if (jobs.n == 0 || total_tiles == 0) return B200_OK;
  launch_pdl(cast_transpose_multi_kernel, dim3(static_cast<unsigned>(total_tiles)), dim3(256), 0, stream, in_base, reinterpret_cast<bf16*>(out_base), jobs);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_cast_f32_bf16(const float* in, void* out, long long n, void* stream) {
  if (n == 0) return B200_OK;
  B200_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0, "cast: alignment");
  const long long thr = (n + 3) / 4;
  launch_pdl(cast_f32_bf16_kernel, dim3((unsigned)((thr + 255) / 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), in, reinterpret_cast<bf16*>(out), n);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_colsum_blocks(long long M) {
  long long b = (M + 511) / 512;
  const int cap = b200_num_sms() * 2;
  if (b > cap) b = cap;
  return b < 1 ? 1 : static_cast<int>(b);
}

extern "C" int b200_colsum(const void* x, long long ld, long long M, int N, float* out, float* partial, int accumulate, void* stream) {
  B200_REQUIRE(N % 4 == 0 && ld % 2 == 0, "colsum: N %% 4, ld %% 2");
  if (M == 0) return B200_OK;
  B200_REQUIRE(N % 8 == 0 && ld % 8 == 0, "colsum: N and ld must be multiples of 8");
  const int blocks = b200_colsum_blocks(M);
  const long long rpb = (M + blocks - 1) / blocks;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (N <= 128) {
    launch_pdl(colsum_kernel<16>, dim3(dim3(blocks, (N + 127) / 128)), dim3(256), 0, st, reinterpret_cast<const bf16*>(x), partial, M, N, ld, rpb);
  } else {
    launch_pdl(colsum_kernel<32>, dim3(dim3(blocks, (N + 255) / 256)), dim3(256), 0, st, reinterpret_cast<const bf16*>(x), partial, M, N, ld, rpb);
  }
  B200_LAUNCH_CHECK();
  return b200_splitk_reduce(partial, out, N, blocks, accumulate, stream);
}

extern "C" int b200_opt_chunk_elems(void) { return kChunk; }

extern "C" int b200_optimizer_step_sum(int kind, const void* tensors_dev, const void* chunks_dev, int n_chunks, float grad_scale,
                                       int step_offset, int n_src, long long src_stride, long long src_shift, void* stream) {
  B200_REQUIRE(n_src >= 1 && n_src <= 64 && src_stride >= 0 && src_shift >= 0, "optimizer_step: bad gradient sources (n_src %d)", n_src);
  if (n_chunks == 0) return B200_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  auto T = reinterpret_cast<const B200OptTensor*>(tensors_dev);
  auto Ck = reinterpret_cast<const int2*>(chunks_dev);
  const GradSrc gs{n_src, src_stride, src_shift};
  if (kind == B200_OPT_SGD) launch_pdl(sgd_kernel, dim3(n_chunks), dim3(256), 0, st, T, Ck, grad_scale, step_offset, gs);
  else if (kind == B200_OPT_ADAMW) launch_pdl(adamw_kernel, dim3(n_chunks), dim3(256), 0, st, T, Ck, grad_scale, step_offset, gs);
  else return b200_set_error(B200_ERR_INVALID, "optimizer_step: unknown kind %d", kind);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_optimizer_step(int kind, const void* tensors_dev, const void* chunks_dev, int n_chunks, float grad_scale,
                                   int step_offset, void* stream) {
  return b200_optimizer_step_sum(kind, tensors_dev, chunks_dev, n_chunks, grad_scale, step_offset, 1, 0, 0, stream);
}
