// Large-margin head around the margin-logits GEMM (gemm.cu: EpiMargin): row normalisation
// (F.normalize, losses/large_margin.py:71), fused focal/cross-entropy forward + gradient of the logits
// (losses/losses.py:22-28; gamma = 0 is mean CE), and the normalisation backward.
//
//   cos = e^ . w^            (GEMM, operands are unit bf16 rows)          logits = s * margin(cos)
//   nll_b = lse_b - logit_b[label];  loss = mean_b (1 - p_b)^gamma nll_b,  p_b = exp(-nll_b)
//   G_bc = dloss/dcos_bc = f_b / B * (softmax_bc - onehot_bc) * s * (c == label ? dphi/dcos : 1)
//   de^ = G w^,  dw^ = G^T e^   (two more GEMMs),   dx = (dx^ - x^ <x^, dx^>) / |x|   for x in {e, w}
#include "common.cuh"

#include "b200_fe.h"

namespace {

// one warp per row: fp32 [R, E] -> unit bf16 rows (+ 1/max(|x|, eps))
__global__ void __launch_bounds__(256) unit_rows_kernel(const float* __restrict__ x, bf16* __restrict__ out, float* __restrict__ inv_norm,
                                                        long long R, int E, long long ld_out, float eps, int as_f16) {
  pdl_grid_sync();
  const long long row = (1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* xr = x + row * E;
  float ss = 0.f;
  for (int i = lane * 4; i < E; i += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xr + i));
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
  for (int i = lane * 4; i < E; i += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xr + i));
    uint2 o;
    if (as_f16) { o.x = pack_f16(v.x * inv, v.y * inv); o.y = pack_f16(v.z * inv, v.w * inv); }
    else { o.x = pack_bf16(v.x * inv, v.y * inv); o.y = pack_bf16(v.z * inv, v.w * inv); }
    *reinterpret_cast<uint2*>(out + row * ld_out + i) = o;
  }
}

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int w = 0; w < nw; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// one CTA per sample row
__global__ void __launch_bounds__(256) margin_ce_kernel(const float* __restrict__ logits, long long ldl, const long long* __restrict__ label,
                                                        const float* __restrict__ cos_label, int B, int C, float s, float cos_m, float sin_m,
                                                        float th, int kind, int easy, float gamma, float* __restrict__ loss_rows,
                                                        bf16* __restrict__ G, long long ldg, float* __restrict__ rdot) {
  pdl_grid_sync();
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float* lr = logits + 1LL * b * ldl;
  const int lab = static_cast<int>(label[b]);
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, lr[c]);
  mx = block_reduce(mx, red, true);
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += __expf(lr[c] - mx);
  se = block_reduce(se, red, false);
  const float lse = mx + logf(se);
  // a label outside [0, C) (mis-numbered start_class / label_map in a config): the reference's scatter_ / CrossEntropyLoss
  // raise; here the row's loss - and with it the batch mean the trainer reads back - becomes NaN instead of a silent
  // out-of-bounds read (no device synchronisation on the hot path)
  const bool bad_label = lab < 0 || lab >= C;
  const float nll = bad_label ? __int_as_float(0x7fc00000) : lse - lr[lab];
  const float p = __expf(-nll);
  float lossv, f;
  if (gamma == 0.f) { lossv = nll; f = 1.f; }
  else {
    const float omp = fmaxf(1.f - p, 0.f);
    lossv = powf(omp, gamma) * nll;
    f = powf(omp, gamma) + gamma * nll * p * powf(omp, gamma - 1.f);   // d loss / d nll
  }
  if (threadIdx.x == 0) loss_rows[b] = lossv;
  if (G == nullptr) return;
  // d logits -> d cos
  const float cl = bad_label ? 0.f : cos_label[b];
  float dphi = 1.f;
  if (kind == 0) {
    const float sine = sqrtf(fmaxf(1.f - cl * cl, 0.f));
    const bool use_phi = easy ? (cl > 0.f) : (cl > th);
    dphi = use_phi ? (cos_m + cl * sin_m / fmaxf(sine, 1e-6f)) : 1.f;   // sine -> 0: reference blows up to inf/NaN; clamped
  }
  const float k = f * s / B;
  float dot = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float l = lr[c];
    const float sm = __expf(l - lse);
    float g, cosv;
    if (c == lab) { g = k * (sm - 1.f) * dphi; cosv = cl; }
    else { g = k * sm; cosv = l / s; }
    const bf16 gb = __float2bfloat16_rn(g);
    G[1LL * b * ldg + c] = gb;
    dot += __bfloat162float(gb) * cosv;
  }
  dot = block_reduce(dot, red, false);
  if (threadIdx.x == 0) rdot[b] = dot;
}

// cdot[c] = sum_b G[b,c] * cos[b,c].  32 columns x 8 row lanes per CTA: row lane r sums b = r, r + 8, ... in order, a fixed
// smem tree folds the 8 partial sums (deterministic; a thread per column walking all B rows was latency-bound: 115 us at
// B = 256, C = 10,000 with 40 CTAs on 148 SMs)
__global__ void __launch_bounds__(256) margin_coldot_kernel(const bf16* __restrict__ G, long long ldg, const float* __restrict__ logits,
                                                            long long ldl, const long long* __restrict__ label,
                                                            const float* __restrict__ cos_label, int B, int C, float s,
                                                            float* __restrict__ cdot) {
  pdl_grid_sync();
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float acc = 0.f;
  if (c < C) {
    const float inv_s = 1.0f / s;
#pragma unroll 4
    for (int b = ry; b < B; b += 8) {
      const float cosv = (c == static_cast<int>(label[b])) ? cos_label[b] : logits[1LL * b * ldl + c] * inv_s;
      acc = fmaf(__bfloat162float(G[1LL * b * ldg + c]), cosv, acc);
    }
  }
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < C) {
    float a = red[0][cx];
#pragma unroll
    for (int r = 1; r < 8; ++r) a += red[r][cx];
    cdot[c] = a;
  }
}

// Stand-alone FocalLoss.forward (losses/losses.py:22-28): per-row focal / cross-entropy loss of arbitrary logits and,
// optionally, its gradient wrt the logits (d mean-loss / d logit).  One CTA per row.
__global__ void __launch_bounds__(256) focal_rows_kernel(const float* __restrict__ logits, long long ldl, const long long* __restrict__ label,
                                                         int B, int C, float gamma, float* __restrict__ loss_rows,
                                                         float* __restrict__ dlogits, long long ldd) {
  pdl_grid_sync();
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float* lr = logits + 1LL * b * ldl;
  const int lab = static_cast<int>(label[b]);
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, lr[c]);
  mx = block_reduce(mx, red, true);
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += __expf(lr[c] - mx);
  se = block_reduce(se, red, false);
  const float lse = mx + logf(se);
  const bool bad_label = lab < 0 || lab >= C;
  const float nll = bad_label ? __int_as_float(0x7fc00000) : lse - lr[lab];
  const float p = __expf(-nll);
  float lossv, f;
  if (gamma == 0.f) { lossv = nll; f = 1.f; }
  else {
    const float omp = fmaxf(1.f - p, 0.f);
    lossv = powf(omp, gamma) * nll;
    f = powf(omp, gamma) + gamma * nll * p * powf(omp, gamma - 1.f);
  }
  if (threadIdx.x == 0) loss_rows[b] = lossv;
  if (dlogits == nullptr) return;
  const float k = f / B;
  for (int c = threadIdx.x; c < C; c += blockDim.x) dlogits[1LL * b * ldd + c] = k * (__expf(lr[c] - lse) - (c == lab ? 1.f : 0.f));
}

__global__ void mean_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
  pdl_grid_sync();
  __shared__ float red[8];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += x[i];
  a = block_reduce(a, red, false);
  if (threadIdx.x == 0) out[0] = a / n;
}

// dx = scale * inv * (T - x^ * dot), x^ = x * inv.  Warp per row.  Writes fp32 and/or bf16.
__global__ void __launch_bounds__(256) unit_rows_bwd_kernel(const float* __restrict__ T, const float* __restrict__ x,
                                                            const float* __restrict__ inv_norm, const float* __restrict__ dot,
                                                            const float* __restrict__ scale_ptr, float* __restrict__ out_f32,
                                                            bf16* __restrict__ out_bf16, long long R, int E, int accumulate) {
  pdl_grid_sync();
  const long long row = (1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float inv = inv_norm[row], d = dot[row];
  const float sc = scale_ptr ? scale_ptr[0] : 1.f;
  for (int i = lane * 4; i < E; i += 128) {
    const float4 t = *reinterpret_cast<const float4*>(T + row * E + i);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * E + i));
    float4 o;
    o.x = sc * inv * (t.x - v.x * inv * d); o.y = sc * inv * (t.y - v.y * inv * d);
    o.z = sc * inv * (t.z - v.z * inv * d); o.w = sc * inv * (t.w - v.w * inv * d);
    if (out_f32) {
      float4* dst = reinterpret_cast<float4*>(out_f32 + row * E + i);
      if (accumulate) { const float4 c = *dst; o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
      *dst = o;
    }
    if (out_bf16) {
      uint2 u; u.x = pack_bf16(o.x, o.y); u.y = pack_bf16(o.z, o.w);
      *reinterpret_cast<uint2*>(out_bf16 + row * E + i) = u;
    }
  }
}

}  // namespace

extern "C" int b200_unit_rows(const float* x, void* out, float* inv_norm, long long R, int E, long long ld_out, float eps,
                              int as_f16, void* stream) {
  B200_REQUIRE(E % 4 == 0 && ld_out % 4 == 0, "unit_rows: E and ld_out must be multiples of 4");
  if (R == 0) return B200_OK;
  const long long blocks = (R * 32 + 255) / 256;
  launch_pdl(unit_rows_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, reinterpret_cast<bf16*>(out), inv_norm, R, E, ld_out, eps, as_f16);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_margin_ce(const float* logits, long long ldl, const long long* label, const float* cos_label, int B, int C,
                              float s, float m, int kind, int easy_margin, float gamma, float* loss_rows, float* loss_mean,
                              void* G, long long ldg, float* rdot, float* cdot, void* stream) {
  if (B == 0) return B200_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const double md = m, pi = 3.14159265358979323846;
  launch_pdl(margin_ce_kernel, dim3(B), dim3(256), 0, st, logits, ldl, label, cos_label, B, C, s, (float)cos(md), (float)sin(md), (float)cos(pi - md), kind,
                                      easy_margin, gamma, loss_rows, reinterpret_cast<bf16*>(G), ldg, rdot);
  B200_LAUNCH_CHECK();
  if (loss_mean) {
    launch_pdl(mean_kernel, dim3(1), dim3(256), 0, st, loss_rows, B, loss_mean);
    B200_LAUNCH_CHECK();
  }
  if (G != nullptr && cdot != nullptr) {
    launch_pdl(margin_coldot_kernel, dim3((C + 31) / 32), dim3(256), 0, st, reinterpret_cast<const bf16*>(G), ldg, logits, ldl, label, cos_label, B, C, s, cdot);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}

extern "C" int b200_focal_loss(const float* logits, long long ldl, const long long* label, int B, int C, float gamma,
                               float* loss_rows, float* loss_mean, float* dlogits, long long ldd, void* stream) {
  B200_REQUIRE(C > 0 && ldl >= C && (dlogits == nullptr || ldd >= C), "focal_loss: bad shape C=%d ldl=%lld ldd=%lld", C, ldl, ldd);
  if (B == 0) return B200_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  launch_pdl(focal_rows_kernel, dim3(B), dim3(256), 0, st, logits, ldl, label, B, C, gamma, loss_rows, dlogits, ldd);
  B200_LAUNCH_CHECK();
  if (loss_mean) {
    launch_pdl(mean_kernel, dim3(1), dim3(256), 0, st, loss_rows, B, loss_mean);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}

extern "C" int b200_unit_rows_bwd(const float* T, const float* x, const float* inv_norm, const float* dot, const float* scale_dev,
                                  float* out_f32, void* out_bf16, long long R, int E, int accumulate, void* stream) {
  B200_REQUIRE(E % 4 == 0, "unit_rows_bwd: E must be a multiple of 4");
  if (R == 0) return B200_OK;
  const long long blocks = (R * 32 + 255) / 256;
  launch_pdl(unit_rows_bwd_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      T, x, inv_norm, dot, scale_dev, out_f32, reinterpret_cast<bf16*>(out_bf16), R, E, accumulate);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
