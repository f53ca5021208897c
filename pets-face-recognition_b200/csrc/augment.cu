// Training-time image augmentation on the GPU: the torchvision / PIL pipeline of the reference's FE configs
// (configs/dog_fe/fe_dogs_config.py:17-26)
//
//   ToPILImage -> RandomAdjustSharpness(0, p=0.1) -> RandomAutocontrast(p=0.3) -> RandomCrop(220) -> Resize(224) ->
//   RandomRotation(5) -> ToTensor
//
// for a whole uint8 batch resident in HBM, restated at the level of PIL's own integer / float32 arithmetic so that, given the
// same random draws (b200/../data_loading/gpu_augment.py draws them in torchvision's order), the output bytes equal PIL's:
//   * sharpness factor 0 = ImageFilter.SMOOTH, the 3x3 kernel (1 1 1 / 1 5 1 / 1 1 1) / 13 in float32, + 0.5, truncated; the
//     one-pixel border is copied (libImaging/Filter.c: ImagingFilter3x3)
//   * autocontrast = per-channel LUT int(v * 255.0 / (hi - lo) - lo * 255.0 / (hi - lo)) clamped, lo / hi = extreme populated
//     histogram bins of the WHOLE (possibly smoothed) image (ImageOps.autocontrast, cutoff 0), in float64 like Python
//   * crop, then PIL's two-pass bilinear resize: horizontal pass to uint8, vertical pass to uint8, 22-bit fixed-point
//     coefficients (libImaging/Resample.c: precompute_coeffs / normalize_coeffs_8bpc; the table comes from the host)
//   * rotation with the NEAREST filter = PIL's 16.16 fixed-point affine walk (libImaging/Geometry.c: affine_fixed), fill 0
// ToTensor's / 255 is not done here: the backbone's first kernel takes uint8 pixels (b200_patch_gather_image_u8).
//
// HBM-bound byte work: one pass for the autocontrast extrema (only the images that drew it), one gather pass for the output;
// every output byte is computed independently (a thread per output pixel, all three channels), reads hit L1 / L2.
#include "common.cuh"

#include "b200_fe.h"

namespace {

struct AugArgs {
  const uint8_t* in;
  uint8_t* out;
  const B200AugParams* prm;
  const int* coef;                // [S][4]: first source index, three 22-bit fixed-point weights (same table for x and y)
  uint8_t* minmax;                // [B][3][2]
  int B, H, W, crop, S;
};

__device__ __forceinline__ int clip8f(float v) { return v <= 0.0f ? 0 : (v >= 255.0f ? 255 : static_cast<int>(v)); }

// ImageFilter.SMOOTH at (y, x) of one channel plane, PIL's evaluation order (no fused multiply-add: PIL's x86 builds have none)
__device__ __forceinline__ int smooth_at(const uint8_t* __restrict__ pl, int H, int W, int y, int x) {
  if (y == 0 || x == 0 || y == H - 1 || x == W - 1) return pl[y * W + x];
  const float k1 = 1.0f / 13.0f, k5 = 5.0f / 13.0f;
  const uint8_t* r1 = pl + (y + 1) * W + x;
  const uint8_t* r0 = pl + y * W + x;
  const uint8_t* r_1 = pl + (y - 1) * W + x;
  float ss = 0.5f;
  ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn(static_cast<float>(r1[-1]), k1), __fmul_rn(static_cast<float>(r1[0]), k1)), __fmul_rn(static_cast<float>(r1[1]), k1)));
  ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn(static_cast<float>(r0[-1]), k1), __fmul_rn(static_cast<float>(r0[0]), k5)), __fmul_rn(static_cast<float>(r0[1]), k1)));
  ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn(static_cast<float>(r_1[-1]), k1), __fmul_rn(static_cast<float>(r_1[0]), k1)), __fmul_rn(static_cast<float>(r_1[1]), k1)));
  return clip8f(ss);
}

// lo / hi of every (image, channel) that drew autocontrast: one CTA per plane
__global__ void __launch_bounds__(256) aug_minmax_kernel(const AugArgs a) {
  pdl_grid_sync();
  const int b = blockIdx.x / 3;
  const B200AugParams p = a.prm[b];
  if (!p.autocontrast) return;
  const uint8_t* pl = a.in + 1LL * blockIdx.x * a.H * a.W;
  int lo = 255, hi = 0;
  for (int i = threadIdx.x; i < a.H * a.W; i += blockDim.x) {
    const int v = p.sharpen ? smooth_at(pl, a.H, a.W, i / a.W, i % a.W) : pl[i];
    lo = min(lo, v);
    hi = max(hi, v);
  }
  __shared__ int slo[8], shi[8];
  for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { lo = min(lo, slo[w]); hi = max(hi, shi[w]); }
    a.minmax[2 * blockIdx.x] = static_cast<uint8_t>(lo);
    a.minmax[2 * blockIdx.x + 1] = static_cast<uint8_t>(hi);
  }
}

// the pixel of channel plane `pl` after sharpness + autocontrast (what the crop / resize read)
__device__ __forceinline__ int source_px(const uint8_t* __restrict__ pl, int H, int W, int y, int x, int sharpen, int ac, double scale, double offset) {
  int v = sharpen ? smooth_at(pl, H, W, y, x) : pl[y * W + x];
  if (ac) {
    const int t = static_cast<int>(__dadd_rn(__dmul_rn(static_cast<double>(v), scale), offset));     // Python: int(ix * scale + offset)
    v = t < 0 ? 0 : (t > 255 ? 255 : t);
  }
  return v;
}

__global__ void __launch_bounds__(256) aug_apply_kernel(const AugArgs a) {
  pdl_grid_sync();
  const int S = a.S;
  const long long idx = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 1LL * a.B * S * S) return;
  const int b = static_cast<int>(idx / (S * S));
  const int rem = static_cast<int>(idx - 1LL * b * S * S);
  const int oy = rem / S, ox = rem - oy * S;
  const B200AugParams p = a.prm[b];
  // rotation, NEAREST: PIL walks xx = a2 + a0 x (+ a1 per row) in 16.16 fixed point and reads pixel (xx >> 16, yy >> 16)
  const int xx = p.rot[2] + p.rot[1] * oy + p.rot[0] * ox;
  const int yy = p.rot[5] + p.rot[4] * oy + p.rot[3] * ox;
  const int rx = xx >> 16, ry = yy >> 16;
  uint8_t* o = a.out + (1LL * b * 3) * S * S + oy * S + ox;
  if (rx < 0 || rx >= S || ry < 0 || ry >= S) {
    o[0] = 0; o[1LL * S * S] = 0; o[2LL * S * S] = 0;
    return;
  }
  // pixel (ry, rx) of the resized image: vertical pass over the rows of the horizontal pass
  const int4 cx = *reinterpret_cast<const int4*>(a.coef + 4 * rx);
  const int4 cy = *reinterpret_cast<const int4*>(a.coef + 4 * ry);
  const int kx[3] = {cx.y, cx.z, cx.w}, ky[3] = {cy.y, cy.z, cy.w};
  for (int c = 0; c < 3; ++c) {
    const uint8_t* pl = a.in + (1LL * b * 3 + c) * a.H * a.W;
    double scale = 1.0, offset = 0.0;
    int ac = 0;
    if (p.autocontrast) {
      const int lo = a.minmax[2 * (b * 3 + c)], hi = a.minmax[2 * (b * 3 + c) + 1];
      if (hi > lo) {
        ac = 1;
        scale = 255.0 / static_cast<double>(hi - lo);
        offset = __dmul_rn(-static_cast<double>(lo), scale);
      }
    }
    int acc_v = 1 << 21;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int sy = cy.x + j;
      if (ky[j] == 0 || sy >= a.crop) continue;
      int acc_h = 1 << 21;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int sx = cx.x + i;
        if (kx[i] == 0 || sx >= a.crop) continue;
        acc_h += kx[i] * source_px(pl, a.H, a.W, p.crop_y + sy, p.crop_x + sx, p.sharpen, ac, scale, offset);
      }
      const int h = min(255, max(0, acc_h >> 22));
      acc_v += ky[j] * h;
    }
    o[1LL * c * S * S] = static_cast<uint8_t>(min(255, max(0, acc_v >> 22)));
  }
}

}  // namespace

extern "C" int b200_augment_train(const unsigned char* in, unsigned char* out, const B200AugParams* params_dev, const int* coef_dev, int B,
                                  int H, int W, int crop, int S, unsigned char* minmax_scratch, void* stream) {
  B200_REQUIRE(B >= 0 && H >= 3 && W >= 3 && crop >= 1 && crop <= H && crop <= W && S >= 1, "augment_train: bad geometry H=%d W=%d crop=%d S=%d", H, W, crop, S);
  B200_REQUIRE(1LL * B * S * S < (1LL << 40) && 1LL * H * W < (1LL << 31), "augment_train: image too large");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(coef_dev) & 15) == 0, "augment_train: coefficient table must be 16-B aligned");
  if (B == 0) return B200_OK;
  AugArgs a{in, out, params_dev, coef_dev, minmax_scratch, B, H, W, crop, S};
  auto st = reinterpret_cast<cudaStream_t>(stream);
  launch_pdl(aug_minmax_kernel, dim3(3 * B), dim3(256), 0, st, a);
  B200_LAUNCH_CHECK();
  const long long n = 1LL * B * S * S;
  launch_pdl(aug_apply_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
