// HBM-bound kernels of the convolutional FE backbone (torchvision.models.resnet50 with fc -> 512, the model the reference's
// FE configs ship: configs/dog_fe/fe_dogs_config.py:96-109).  The convolutions themselves are tcgen05 GEMMs (gemm.cu:
// b200_gemm_tn for 1x1, b200_gemm_taps for 3x3 - implicit, no im2col matrix); this file holds what sits between them.
//
// Activation layout: bf16 NHWC on a PADDED grid.  An image of H x W pixels occupies (H + 2) x (W + 2) consecutive rows of a
// [rows, C] matrix - one ring of zero rows around it - so that (a) a 1x1 convolution is a plain GEMM over all rows, (b) a 3x3
// convolution is nine row-shifted GEMM contributions (shift = dy * (W + 2) + dx; the ring supplies the zero padding), and
// (c) every kernel here touches whole rows, 16 B per thread, fully coalesced.  Ring rows hold garbage after a convolution;
// the BatchNorm kernels skip them in the statistics and write them back as zeros.
//
//   bn_stats / bn_finalize     per-channel batch mean, biased variance -> scale / shift, running statistics (momentum 0.1,
//                              unbiased variance) as nn.BatchNorm2d does in training mode
//   bn_apply                   y = [relu](x * scale + shift [+ residual]) on interior rows, 0 on the ring
//   bn_bwd_stats / _apply      dz = dy * [y > 0];  sum dz, sum dz * xhat;  dx = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat))
//   stem_im2col                7x7 stride-2 pad-3 patches of the uint8 / float image -> [B * 112 * 112, 160] bf16 (147 + zero pad)
//   stem_pool fwd / bwd        BatchNorm + ReLU + MaxPool2d(3, 2, 1) fused; the arg-max tap is kept for the backward
//   grid_subsample / upsample  stride-2 sampling between grids (the stride-2 convolutions are computed at stride 1 or on
//                              the sampled rows) and its adjoint
//   grid_avgpool fwd / bwd     AdaptiveAvgPool2d(1) over the interior of the last grid
#include "common.cuh"

#include "b200_fe.h"

namespace {

struct Grid {        // H = 0: a plain [rows, C] matrix (every row is interior)
  int H, W, Wp, P;   // Wp = W + 2, P = (H + 2) * (W + 2)
};
inline Grid make_grid(int H, int W) { return Grid{H, W, W + 2, (H + 2) * (W + 2)}; }
__device__ __forceinline__ bool interior(const Grid& g, long long row) {
  if (g.H == 0) return true;
  const int q = static_cast<int>(row % g.P);
  const int y = q / g.Wp, x = q - y * g.Wp;
  return y >= 1 && y <= g.H && x >= 1 && x <= g.W;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* x) {
  float2 f;
  f = unpack_bf16(u.x); x[0] = f.x; x[1] = f.y;
  f = unpack_bf16(u.y); x[2] = f.x; x[3] = f.y;
  f = unpack_bf16(u.z); x[4] = f.x; x[5] = f.y;
  f = unpack_bf16(u.w); x[6] = f.x; x[7] = f.y;
}
__device__ __forceinline__ uint4 pack8(const float* x) {
  uint4 u;
  u.x = pack_bf16(x[0], x[1]); u.y = pack_bf16(x[2], x[3]); u.z = pack_bf16(x[4], x[5]); u.w = pack_bf16(x[6], x[7]);
  return u;
}
__device__ __forceinline__ uint4 ldg16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void ld8f(const float* p, float* x) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}

constexpr int kStatThreads = 256;

// Two per-channel sums over the interior rows of a row range per CTA.  MODE 0: sum x, sum x^2 (forward statistics);
// MODE 1: sum dz, sum dz * xhat with dz = dy * [y > 0] (y null: no ReLU), xhat = (x - mean) * rstd.
// Thread = (row lane, 8-channel chunk); C / 8 is a power of two <= 256.  partial: [blocks][2][C].
template <int MODE>
__global__ void __launch_bounds__(kStatThreads) bn_sums_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const float* __restrict__ relu_scale, const float* __restrict__ relu_shift,
                                                               long long rows, int C, int cpr_log2, Grid g, float* __restrict__ partial) {
  pdl_grid_sync();
  __shared__ float red[kStatThreads * 16];
  const int cpr = 1 << cpr_log2;
  const int ch = threadIdx.x & (cpr - 1), rl = threadIdx.x >> cpr_log2, lanes = kStatThreads >> cpr_log2;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = per * blockIdx.x, r1 = min(rows, r0 + per);
  float a[8], b[8], mu[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 0.f; b[i] = 0.f; mu[i] = 0.f; rs[i] = 1.f; }
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = 0.f; sh[i] = 0.f; }
  if (MODE == 1) { ld8f(mean + ch * 8, mu); ld8f(rstd + ch * 8, rs); }
  if (MODE == 1 && relu_scale != nullptr) { ld8f(relu_scale + ch * 8, sc); ld8f(relu_shift + ch * 8, sh); }
  // four rows per iteration, every load issued before the arithmetic: the kernel is a pure stream and needs the loads in flight
  constexpr int U = 4;
  for (long long rb = r0 + rl; rb < r1; rb += 1LL * lanes * U) {
    uint4 rx[U], rd[U], ry[U];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rb + 1LL * u * lanes;
      live[u] = r < r1 && interior(g, r);
      if (live[u]) {
        rx[u] = ldg16(x + r * C + ch * 8);
        if (MODE == 1) {
          rd[u] = ldg16(dy + r * C + ch * 8);
          if (y != nullptr) ry[u] = ldg16(y + r * C + ch * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!live[u]) continue;
      float xv[8];
      unpack8(rx[u], xv);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += xv[i]; b[i] = fmaf(xv[i], xv[i], b[i]); }
      } else {
        float dv[8], yv[8];
        unpack8(rd[u], dv);
        if (y != nullptr) {
          unpack8(ry[u], yv);
#pragma unroll
          for (int i = 0; i < 8; ++i) dv[i] = yv[i] > 0.f ? dv[i] : 0.f;
        } else if (relu_scale != nullptr) {      // the ReLU mask from the BatchNorm input itself: y = relu(x * scale + shift)
#pragma unroll
          for (int i = 0; i < 8; ++i) dv[i] = fmaf(xv[i], sc[i], sh[i]) > 0.f ? dv[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += dv[i]; b[i] = fmaf(dv[i], (xv[i] - mu[i]) * rs[i], b[i]); }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[threadIdx.x * 16 + i] = a[i]; red[threadIdx.x * 16 + 8 + i] = b[i]; }
  __syncthreads();
  if (rl == 0) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] += red[((l << cpr_log2) + ch) * 16 + i]; b[i] += red[((l << cpr_log2) + ch) * 16 + 8 + i]; }
    float* o = partial + 2LL * blockIdx.x * C;
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[ch * 8 + i] = a[i]; o[C + ch * 8 + i] = b[i]; }
  }
}

// fold the CTA partials in double, one warp per channel (fixed order: deterministic); MODE 0 also turns them into the
// normalisation constants
template <int MODE>
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partial, int blocks, int C, double count,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
                                                          float* running_var, float momentum, float eps,
                                                          float* __restrict__ out /* MODE 0: [4][C] scale, shift, mean, rstd; MODE 1: [2][C] sums */) {
  pdl_grid_sync();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int b = lane; b < blocks; b += 32) { s0 += partial[2LL * b * C + c]; s1 += partial[2LL * b * C + C + c]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  if (lane != 0) return;
  if (MODE == 1) { out[c] = static_cast<float>(s0); out[C + c] = static_cast<float>(s1); return; }
  const double mean = s0 / count;
  double var = s1 / count - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float scale = gamma[c] * rstd;
  out[c] = scale;
  out[C + c] = beta[c] - static_cast<float>(mean) * scale;
  out[2 * C + c] = static_cast<float>(mean);
  out[3 * C + c] = rstd;
  if (running_mean != nullptr) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                                       const bf16* __restrict__ res, int relu, long long rows, int C, int cpr_log2, Grid g,
                                                       bf16* __restrict__ y) {
  pdl_grid_sync();
  const long long t = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = t >> cpr_log2;
  if (row >= rows) return;
  const int ch = static_cast<int>(t & ((1 << cpr_log2) - 1));
  const long long off = row * C + ch * 8;
  if (!interior(g, row)) { *reinterpret_cast<uint4*>(y + off) = make_uint4(0u, 0u, 0u, 0u); return; }
  float v[8], sc[8], sh[8];
  unpack8(ldg16(x + off), v);
  ld8f(scale + ch * 8, sc);
  ld8f(shift + ch * 8, sh);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc[i], sh[i]);
  if (res != nullptr) {
    float r[8];
    unpack8(ldg16(res + off), r);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += r[i];
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  *reinterpret_cast<uint4*>(y + off) = pack8(v);
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y, const bf16* __restrict__ x,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ sums, float inv_count,
                                                           const float* __restrict__ relu_scale, const float* __restrict__ relu_shift,
                                                           long long rows, int C, int cpr_log2, Grid g, bf16* __restrict__ dx, bf16* dz_out) {
  pdl_grid_sync();
  const long long t = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = t >> cpr_log2;
  if (row >= rows) return;
  const int ch = static_cast<int>(t & ((1 << cpr_log2) - 1));
  const long long off = row * C + ch * 8;
  if (!interior(g, row)) {
    *reinterpret_cast<uint4*>(dx + off) = make_uint4(0u, 0u, 0u, 0u);
    if (dz_out != nullptr) *reinterpret_cast<uint4*>(dz_out + off) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  float dv[8], xv[8], mu[8], rs[8], gm[8], s0[8], s1[8], o[8];
  unpack8(ldg16(dy + off), dv);
  unpack8(ldg16(x + off), xv);
  if (y != nullptr) {
    float yv[8];
    unpack8(ldg16(y + off), yv);
#pragma unroll
    for (int i = 0; i < 8; ++i) dv[i] = yv[i] > 0.f ? dv[i] : 0.f;
  } else if (relu_scale != nullptr) {
    float sc[8], sh[8];
    ld8f(relu_scale + ch * 8, sc); ld8f(relu_shift + ch * 8, sh);
#pragma unroll
    for (int i = 0; i < 8; ++i) dv[i] = fmaf(xv[i], sc[i], sh[i]) > 0.f ? dv[i] : 0.f;
  }
  ld8f(mean + ch * 8, mu); ld8f(rstd + ch * 8, rs); ld8f(gamma + ch * 8, gm);
  ld8f(sums + ch * 8, s0); ld8f(sums + C + ch * 8, s1);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float xh = (xv[i] - mu[i]) * rs[i];
    o[i] = gm[i] * rs[i] * (dv[i] - s0[i] * inv_count - xh * s1[i] * inv_count);
  }
  *reinterpret_cast<uint4*>(dx + off) = pack8(o);
  if (dz_out != nullptr) *reinterpret_cast<uint4*>(dz_out + off) = pack8(dv);
}

// ------------------------------------------------------------------------------------------------ stem
// cols[(b, oy, ox), (r * 7 + s) * 3 + c] = img[b, c, 2 oy - 3 + r, 2 ox - 3 + s] (0 outside), columns 147..159 zero.
// One CTA per output row (b, oy): the 3 x 7 input row segments it needs are staged in shared memory with coalesced loads
// (zero-padded by 3 pixels on either side), then every thread assembles 16-B chunks of the patch matrix from there and
// writes them coalesced - the kernel is bound by the 1 GB of patches it writes at batch 256, not by byte gathers.
template <bool U8>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const void* __restrict__ img, int B, int H, int W, int OH, int OW, bf16* __restrict__ cols) {
  pdl_grid_sync();
  extern __shared__ float stem_tile[];               // [3][7][W + 6], already scaled to [0, 1]
  const int b = blockIdx.x / OH, oy = blockIdx.x - b * OH;
  const int Wt = W + 6;
  for (int idx = threadIdx.x; idx < 21 * Wt; idx += blockDim.x) {
    const int cr = idx / Wt, xx = idx - cr * Wt;
    const int c = cr / 7, r = cr - c * 7;
    const int iy = 2 * oy - 3 + r, ix = xx - 3;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const long long g = ((1LL * b * 3 + c) * H + iy) * W + ix;
      v = U8 ? static_cast<float>(__ldg(static_cast<const uint8_t*>(img) + g)) * (1.0f / 255.0f) : __ldg(static_cast<const float*>(img) + g);
    }
    stem_tile[idx] = v;
  }
  __syncthreads();
  bf16* out = cols + (1LL * b * OH + oy) * OW * 160;
  for (int j = threadIdx.x; j < OW * 20; j += blockDim.x) {
    const int ox = j / 20, chunk = j - ox * 20;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = chunk * 8 + i;
      float val = 0.f;
      if (k < 147) {
        const int r = k / 21, rem = k - r * 21, s2 = rem / 3, c = rem - s2 * 3;
        val = stem_tile[(c * 7 + r) * Wt + 2 * ox + s2];
      }
      v[i] = val;
    }
    *reinterpret_cast<uint4*>(out + 8LL * j) = pack8(v);
  }
}

// y[(b, py, px)] on the padded (OH/2 + 2) x (OW/2 + 2) grid = max over the 3x3 stride-2 pad-1 window of relu(a * scale + shift);
// tap = index (r * 3 + s) of the first maximum in scan order (torch's choice)
__global__ void __launch_bounds__(256) stem_pool_fwd_kernel(const bf16* __restrict__ a, const float* __restrict__ scale, const float* __restrict__ shift,
                                                            int B, int IH, int IW, int C, bf16* __restrict__ y, uint8_t* __restrict__ tap) {
  pdl_grid_sync();
  const int cpr = C >> 3;
  const int OH = IH / 2, OW = IW / 2, Wp = OW + 2, P = (OH + 2) * Wp;
  const long long t = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = t / cpr;
  if (row >= 1LL * B * P) return;
  const int ch = static_cast<int>(t - row * cpr);
  const int q = static_cast<int>(row % P), b = static_cast<int>(row / P);
  const int py = q / Wp, px = q - py * Wp;
  if (py < 1 || py > OH || px < 1 || px > OW) { *reinterpret_cast<uint4*>(y + row * C + ch * 8) = make_uint4(0u, 0u, 0u, 0u); return; }
  const int oy = py - 1, ox = px - 1;
  float sc[8], sh[8], best[8];
  int bt[8];
  ld8f(scale + ch * 8, sc);
  ld8f(shift + ch * 8, sh);
#pragma unroll
  for (int i = 0; i < 8; ++i) { best[i] = -INFINITY; bt[i] = 0; }
  for (int r = 0; r < 3; ++r) {
    const int iy = 2 * oy - 1 + r;
    if (iy < 0 || iy >= IH) continue;
    for (int s = 0; s < 3; ++s) {
      const int ix = 2 * ox - 1 + s;
      if (ix < 0 || ix >= IW) continue;
      float v[8];
      unpack8(ldg16(a + ((1LL * b * IH + iy) * IW + ix) * C + ch * 8), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float z = fmaxf(fmaf(v[i], sc[i], sh[i]), 0.f);
        if (z > best[i]) { best[i] = z; bt[i] = r * 3 + s; }
      }
    }
  }
  *reinterpret_cast<uint4*>(y + row * C + ch * 8) = pack8(best);
  const long long orow = (1LL * b * OH + oy) * OW + ox;
  uint2 tp;
  tp.x = bt[0] | (bt[1] << 8) | (bt[2] << 16) | (bt[3] << 24);
  tp.y = bt[4] | (bt[5] << 8) | (bt[6] << 16) | (bt[7] << 24);
  *reinterpret_cast<uint2*>(tap + orow * C + ch * 8) = tp;
}

// dz[(b, iy, ix)] = [a * scale + shift > 0] * sum over the pooling windows whose arg-max is this pixel of dy
__global__ void __launch_bounds__(256) stem_pool_bwd_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ tap, const bf16* __restrict__ a,
                                                            const float* __restrict__ scale, const float* __restrict__ shift, int B, int IH, int IW,
                                                            int C, bf16* __restrict__ dz) {
  pdl_grid_sync();
  const int cpr = C >> 3;
  const int OH = IH / 2, OW = IW / 2, Wp = OW + 2, P = (OH + 2) * Wp;
  const long long t = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = t / cpr;
  if (row >= 1LL * B * IH * IW) return;
  const int ch = static_cast<int>(t - row * cpr);
  const int ix = static_cast<int>(row % IW), iy = static_cast<int>((row / IW) % IH), b = static_cast<int>(row / (1LL * IW * IH));
  float acc[8], v[8], sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  // windows (oy, ox) with 2 o - 1 + r = i: even i -> (i / 2, r = 1); odd i -> ((i - 1) / 2, r = 2) and ((i + 1) / 2, r = 0)
  const int ny = (iy & 1) ? 2 : 1, nx = (ix & 1) ? 2 : 1;
  for (int jy = 0; jy < ny; ++jy) {
    const int oy = (iy & 1) ? (iy - 1) / 2 + jy : iy / 2;
    const int r = (iy & 1) ? (jy == 0 ? 2 : 0) : 1;
    if (oy >= OH) continue;
    for (int jx = 0; jx < nx; ++jx) {
      const int ox = (ix & 1) ? (ix - 1) / 2 + jx : ix / 2;
      const int s = (ix & 1) ? (jx == 0 ? 2 : 0) : 1;
      if (ox >= OW) continue;
      const uint2 tp = __ldg(reinterpret_cast<const uint2*>(tap + ((1LL * b * OH + oy) * OW + ox) * C + ch * 8));
      float d[8];
      unpack8(ldg16(dy + (1LL * b * P + (oy + 1) * Wp + ox + 1) * C + ch * 8), d);
      const int want = r * 3 + s;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ti = ((i < 4 ? tp.x : tp.y) >> (8 * (i & 3))) & 0xff;
        if (ti == want) acc[i] += d[i];
      }
    }
  }
  unpack8(ldg16(a + row * C + ch * 8), v);
  ld8f(scale + ch * 8, sc);
  ld8f(shift + ch * 8, sh);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = fmaf(v[i], sc[i], sh[i]) > 0.f ? acc[i] : 0.f;
  *reinterpret_cast<uint4*>(dz + row * C + ch * 8) = pack8(acc);
}

// ------------------------------------------------------------------------------------------------ grids
// DOWN: out (H/2 x W/2 grid) interior (oy, ox) = in (H x W grid) interior (2 oy, 2 ox).  UP (the adjoint): in-grid gradient =
// out-grid gradient at the sampled pixels, zero elsewhere.  `rows` counts the rows of the tensor being written.
template <bool DOWN>
__global__ void __launch_bounds__(256) grid_sample_kernel(const bf16* __restrict__ src, int B, int H, int W, int C, bf16* __restrict__ dst) {
  pdl_grid_sync();
  const int cpr = C >> 3;
  const int h = H / 2, w = W / 2;
  const int Wp = W + 2, P = (H + 2) * Wp, wp = w + 2, pp = (h + 2) * wp;
  const long long t = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = t / cpr;
  const int ch = static_cast<int>(t - row * cpr);
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (DOWN) {
    if (row >= 1LL * B * pp) return;
    const int q = static_cast<int>(row % pp), b = static_cast<int>(row / pp);
    const int py = q / wp, px = q - py * wp;
    if (py >= 1 && py <= h && px >= 1 && px <= w) v = ldg16(src + (1LL * b * P + (2 * (py - 1) + 1) * Wp + 2 * (px - 1) + 1) * C + ch * 8);
  } else {
    if (row >= 1LL * B * P) return;
    const int q = static_cast<int>(row % P), b = static_cast<int>(row / P);
    const int py = q / Wp, px = q - py * Wp;
    if (py >= 1 && py <= H && px >= 1 && px <= W && ((py - 1) & 1) == 0 && ((px - 1) & 1) == 0)
      v = ldg16(src + (1LL * b * pp + ((py - 1) / 2 + 1) * wp + (px - 1) / 2 + 1) * C + ch * 8);
  }
  *reinterpret_cast<uint4*>(dst + row * C + ch * 8) = v;
}

// Patches of a stride-2 3x3 convolution, for its weight gradient: cols[(b, py, px) on the (H/2, W/2) grid, t * C + c] =
// x[(b, 2 (py - 1) + 1 + r - 1, 2 (px - 1) + 1 + s - 1) on the (H, W) grid, c], t = r * 3 + s; ring rows of the small grid are zero.
// (The forward and the data gradient of these three convolutions run on the fine grid; only the weight gradient would pay four
// times the contraction length for it.)
__global__ void __launch_bounds__(256) grid_patches_s2_kernel(const bf16* __restrict__ x, int B, int H, int W, int C, bf16* __restrict__ cols) {
  pdl_grid_sync();
  const int cpr = C >> 3;
  const int h = H / 2, w = W / 2;
  const int Wp = W + 2, P = (H + 2) * Wp, wp = w + 2, pp = (h + 2) * wp;
  const long long t = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
  const long long chunk = t / cpr;               // (row of the small grid, tap)
  const int ch = static_cast<int>(t - chunk * cpr);
  const long long row = chunk / 9;
  const int tap = static_cast<int>(chunk - row * 9);
  if (row >= 1LL * B * pp) return;
  const int q = static_cast<int>(row % pp), b = static_cast<int>(row / pp);
  const int py = q / wp, px = q - py * wp;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (py >= 1 && py <= h && px >= 1 && px <= w) {
    const int r = tap / 3, sft = tap - r * 3;
    v = ldg16(x + (1LL * b * P + (2 * (py - 1) + r) * Wp + 2 * (px - 1) + sft) * C + ch * 8);      // padded coordinates: +1 - 1
  }
  *reinterpret_cast<uint4*>(cols + (row * 9 + tap) * C + ch * 8) = v;
}

__global__ void __launch_bounds__(256) grid_avgpool_fwd_kernel(const bf16* __restrict__ x, int B, int H, int W, int C, bf16* __restrict__ out) {
  pdl_grid_sync();
  const int cpr = C >> 3;
  const long long t = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 1LL * B * cpr) return;
  const int b = static_cast<int>(t / cpr), ch = static_cast<int>(t - 1LL * b * cpr);
  const int Wp = W + 2, P = (H + 2) * Wp;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int y = 1; y <= H; ++y)
    for (int xx = 1; xx <= W; ++xx) {
      float v[8];
      unpack8(ldg16(x + (1LL * b * P + y * Wp + xx) * C + ch * 8), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  const float inv = 1.0f / static_cast<float>(H * W);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] *= inv;
  *reinterpret_cast<uint4*>(out + 1LL * b * C + ch * 8) = pack8(acc);
}

__global__ void __launch_bounds__(256) grid_avgpool_bwd_kernel(const bf16* __restrict__ dout, int B, int H, int W, int C, bf16* __restrict__ dx) {
  pdl_grid_sync();
  const int cpr = C >> 3;
  const int Wp = W + 2, P = (H + 2) * Wp;
  const long long t = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = t / cpr;
  if (row >= 1LL * B * P) return;
  const int ch = static_cast<int>(t - row * cpr);
  const int q = static_cast<int>(row % P), b = static_cast<int>(row / P);
  const int py = q / Wp, px = q - py * Wp;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (py >= 1 && py <= H && px >= 1 && px <= W) {
    float v[8];
    unpack8(ldg16(dout + 1LL * b * C + ch * 8), v);
    const float inv = 1.0f / static_cast<float>(H * W);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= inv;
    o = pack8(v);
  }
  *reinterpret_cast<uint4*>(dx + row * C + ch * 8) = o;
}

int log2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}
unsigned blocks_for(long long threads) { return static_cast<unsigned>((threads + 255) / 256); }

}  // namespace

#define REQ_C(C) B200_REQUIRE((C) >= 8 && (C) <= 2048 && log2_exact((C) / 8) >= 0 && (C) % 8 == 0, "convnet: channels must be 8 * 2^k <= 2048 (got %d)", (C))
#define REQ_ALIGN(p) B200_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "convnet: tensors must be 16-B aligned")

extern "C" int b200_bn_stats_blocks(long long rows) {
  const long long want = (rows + 255) / 256;
  const long long cap = 4LL * b200_num_sms();
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

// Training-mode BatchNorm statistics of x [rows, C] over the interior rows of an (H, W) grid (H = 0: all rows), count = their
// number.  out [4][C] = scale, shift, mean, rstd; the running statistics (may be null) are updated as nn.BatchNorm2d does.
// scratch: [b200_bn_stats_blocks(rows)][2][C] floats.
extern "C" int b200_bn_stats(const void* x, long long rows, int C, int H, int W, double count, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, float momentum, float eps, float* out, float* scratch, void* stream) {
  REQ_C(C); REQ_ALIGN(x);
  B200_REQUIRE(rows > 0 && count > 0, "bn_stats: empty input");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = b200_bn_stats_blocks(rows);
  const Grid g = H > 0 ? make_grid(H, W) : Grid{0, 0, 0, 1};
  launch_pdl(bn_sums_kernel<0>, dim3(blocks), dim3(kStatThreads), 0, st, static_cast<const bf16*>(x), static_cast<const bf16*>(nullptr),
             static_cast<const bf16*>(nullptr), static_cast<const float*>(nullptr), static_cast<const float*>(nullptr), static_cast<const float*>(nullptr),
             static_cast<const float*>(nullptr), rows, C, log2_exact(C / 8), g, scratch);
  B200_LAUNCH_CHECK();
  launch_pdl(bn_finalize_kernel<0>, dim3((C + 7) / 8), dim3(256), 0, st, static_cast<const float*>(scratch), blocks, C, count, gamma, beta,
             running_mean, running_var, momentum, eps, out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_bn_apply(const void* x, const float* scale, const float* shift, const void* residual, int relu, long long rows, int C, int H,
                             int W, void* y, void* stream) {
  REQ_C(C); REQ_ALIGN(x); REQ_ALIGN(y); REQ_ALIGN(residual);
  const Grid g = H > 0 ? make_grid(H, W) : Grid{0, 0, 0, 1};
  launch_pdl(bn_apply_kernel, dim3(blocks_for(rows * (C / 8))), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), static_cast<const bf16*>(x),
             scale, shift, static_cast<const bf16*>(residual), relu, rows, C, log2_exact(C / 8), g, static_cast<bf16*>(y));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// BatchNorm (+ ReLU when y is given, or relu_from_x: y = relu(x * scale + shift) with no residual, mask recomputed) backward.  stats = the [4][C] block of b200_bn_stats (mean / rstd are read), sums [2][C]
// receives dbeta = sum dz and dgamma = sum dz * xhat; dx [rows, C]; dz_out (may be null) = dy * [y > 0], the gradient that
// continues along a residual connection.  count = number of interior rows the statistics were taken over, or 0 for frozen
// statistics (eval-mode BatchNorm: mean / rstd are constants, dx = gamma * rstd * dz).  scratch as for b200_bn_stats.
extern "C" int b200_bn_backward(const void* dy, const void* y, int relu_from_x, const void* x, const float* stats, const float* gamma, long long rows,
                                int C, int H, int W, double count, void* dx, void* dz_out, float* sums, float* scratch, void* stream) {
  REQ_C(C); REQ_ALIGN(dy); REQ_ALIGN(y); REQ_ALIGN(x); REQ_ALIGN(dx); REQ_ALIGN(dz_out);
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = b200_bn_stats_blocks(rows);
  const Grid g = H > 0 ? make_grid(H, W) : Grid{0, 0, 0, 1};
  const float* mean = stats + 2 * C;
  const float* rstd = stats + 3 * C;
  B200_REQUIRE(!(relu_from_x && y != nullptr), "bn_backward: the ReLU mask comes from y or from x, not both");
  const float* rsc = relu_from_x ? stats : nullptr;          // y = relu(x * scale + shift) without a residual: the mask needs no y
  const float* rsh = relu_from_x ? stats + C : nullptr;
  launch_pdl(bn_sums_kernel<1>, dim3(blocks), dim3(kStatThreads), 0, st, static_cast<const bf16*>(x), static_cast<const bf16*>(dy),
             static_cast<const bf16*>(y), mean, rstd, rsc, rsh, rows, C, log2_exact(C / 8), g, scratch);
  B200_LAUNCH_CHECK();
  launch_pdl(bn_finalize_kernel<1>, dim3((C + 7) / 8), dim3(256), 0, st, static_cast<const float*>(scratch), blocks, C, count,
             static_cast<const float*>(nullptr), static_cast<const float*>(nullptr), static_cast<float*>(nullptr), static_cast<float*>(nullptr), 0.f, 0.f, sums);
  B200_LAUNCH_CHECK();
  launch_pdl(bn_bwd_apply_kernel, dim3(blocks_for(rows * (C / 8))), dim3(256), 0, st, static_cast<const bf16*>(dy), static_cast<const bf16*>(y),
             static_cast<const bf16*>(x), mean, rstd, gamma, static_cast<const float*>(sums), static_cast<float>(count > 0 ? 1.0 / count : 0.0), rsc, rsh, rows, C,
             log2_exact(C / 8), g, static_cast<bf16*>(dx), static_cast<bf16*>(dz_out));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_stem_im2col(const void* img, int is_u8, int B, int H, int W, void* cols, void* stream) {
  B200_REQUIRE(B > 0 && H % 2 == 0 && W % 2 == 0 && H >= 8 && W >= 8, "stem_im2col: even image sides wanted (got %d x %d)", H, W);
  REQ_ALIGN(cols);
  const int OH = H / 2, OW = W / 2;
  const size_t smem = sizeof(float) * 21 * (W + 6);
  B200_REQUIRE(smem <= 48 * 1024 && 1LL * B * OH < (1LL << 31), "stem_im2col: image too wide (%d) or batch too large", W);
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (is_u8) launch_pdl(stem_im2col_kernel<true>, dim3(static_cast<unsigned>(B * OH)), dim3(256), smem, st, img, B, H, W, OH, OW, static_cast<bf16*>(cols));
  else launch_pdl(stem_im2col_kernel<false>, dim3(static_cast<unsigned>(B * OH)), dim3(256), smem, st, img, B, H, W, OH, OW, static_cast<bf16*>(cols));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// a: [B * IH * IW, C] convolution output; y: padded (IH / 2, IW / 2) grid; tap: [B * IH/2 * IW/2, C] bytes
extern "C" int b200_stem_pool_fwd(const void* a, const float* scale, const float* shift, int B, int IH, int IW, int C, void* y, void* tap,
                                  void* stream) {
  REQ_C(C); REQ_ALIGN(a); REQ_ALIGN(y);
  B200_REQUIRE(IH % 2 == 0 && IW % 2 == 0 && (reinterpret_cast<uintptr_t>(tap) & 7) == 0, "stem_pool: even sides, 8-B aligned tap tensor");
  const long long threads = 1LL * B * (IH / 2 + 2) * (IW / 2 + 2) * (C / 8);
  launch_pdl(stem_pool_fwd_kernel, dim3(blocks_for(threads)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), static_cast<const bf16*>(a), scale,
             shift, B, IH, IW, C, static_cast<bf16*>(y), static_cast<uint8_t*>(tap));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_stem_pool_bwd(const void* dy, const void* tap, const void* a, const float* scale, const float* shift, int B, int IH, int IW,
                                  int C, void* dz, void* stream) {
  REQ_C(C); REQ_ALIGN(a); REQ_ALIGN(dy); REQ_ALIGN(dz);
  const long long threads = 1LL * B * IH * IW * (C / 8);
  launch_pdl(stem_pool_bwd_kernel, dim3(blocks_for(threads)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), static_cast<const bf16*>(dy),
             static_cast<const uint8_t*>(tap), static_cast<const bf16*>(a), scale, shift, B, IH, IW, C, static_cast<bf16*>(dz));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// down != 0: src on the (H, W) grid -> dst on the (H / 2, W / 2) grid (pixels (2 i, 2 j)); down == 0: the adjoint (dst on (H, W))
extern "C" int b200_grid_sample2(const void* src, int B, int H, int W, int C, int down, void* dst, void* stream) {
  REQ_C(C); REQ_ALIGN(src); REQ_ALIGN(dst);
  B200_REQUIRE(H % 2 == 0 && W % 2 == 0, "grid_sample2: even grid sides wanted");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (down) {
    const long long threads = 1LL * B * (H / 2 + 2) * (W / 2 + 2) * (C / 8);
    launch_pdl(grid_sample_kernel<true>, dim3(blocks_for(threads)), dim3(256), 0, st, static_cast<const bf16*>(src), B, H, W, C, static_cast<bf16*>(dst));
  } else {
    const long long threads = 1LL * B * (H + 2) * (W + 2) * (C / 8);
    launch_pdl(grid_sample_kernel<false>, dim3(blocks_for(threads)), dim3(256), 0, st, static_cast<const bf16*>(src), B, H, W, C, static_cast<bf16*>(dst));
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

/* x on the (H, W) grid -> cols [rows of the (H / 2, W / 2) grid, 9 * C]: the patches a stride-2 3x3 convolution reads */
extern "C" int b200_grid_patches_s2(const void* x, int B, int H, int W, int C, void* cols, void* stream) {
  REQ_C(C); REQ_ALIGN(x); REQ_ALIGN(cols);
  B200_REQUIRE(H % 2 == 0 && W % 2 == 0, "grid_patches_s2: even grid sides wanted");
  const long long threads = 1LL * B * (H / 2 + 2) * (W / 2 + 2) * 9 * (C / 8);
  launch_pdl(grid_patches_s2_kernel, dim3(blocks_for(threads)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), static_cast<const bf16*>(x), B, H, W,
             C, static_cast<bf16*>(cols));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_grid_avgpool(const void* x, int B, int H, int W, int C, int backward, void* out, void* stream) {
  REQ_C(C); REQ_ALIGN(x); REQ_ALIGN(out);
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (!backward) launch_pdl(grid_avgpool_fwd_kernel, dim3(blocks_for(1LL * B * (C / 8))), dim3(256), 0, st, static_cast<const bf16*>(x), B, H, W, C, static_cast<bf16*>(out));
  else launch_pdl(grid_avgpool_bwd_kernel, dim3(blocks_for(1LL * B * (H + 2) * (W + 2) * (C / 8))), dim3(256), 0, st, static_cast<const bf16*>(x), B, H, W, C, static_cast<bf16*>(out));
  B200_LAUNCH_CHECK();
  return B200_OK;
}
