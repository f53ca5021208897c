// Linear-layer GEMMs on the tcgen05 core: forward (bias / GELU / residual epilogues), data-gradient
// (optionally fused with GELU'), weight-gradient (split-K partials + deterministic reduce) and the
// ArcFace logits GEMM with the margin + scale epilogue (losses/large_margin.py:69-84 of the reference).
#include "gemm_core.cuh"

#include <cmath>
#include <cstdlib>
#include <mutex>

#include "b200_fe.h"

// ---------------------------------------------------------------------------------------------
// library-wide error state + device info
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int b200_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
extern "C" const char* b200_last_error(void) { return g_err; }

int b200_wait_ns() {
  static const int ns = [] {
    const char* e = getenv("B200_WAIT_NS");
    const int v = e != nullptr ? atoi(e) : 256;
    return v < 0 ? 0 : v;
  }();
  return ns;
}

int b200_epi16() {
  static const int on = [] {
    const char* e = getenv("B200_EPI16");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
  }();
  return on;
}

namespace gemm {
int b200_cg2() {
  static const int on = [] {
    const char* e = getenv("B200_CG2");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
  }();
  return on;
}
}  // namespace gemm

int b200_reverse_rows() {
  static const int on = [] {
    const char* e = getenv("B200_REVERSE");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  return on;
}

int b200_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

// ---------------------------------------------------------------------------------------------
// profiling hooks: launch counter + optional CUDA-event pairs around every tcgen05 GEMM launch
// ---------------------------------------------------------------------------------------------
#include <atomic>
#include <vector>
static std::atomic<long long> g_launches{0};
static bool g_prof_gemm = false;
bool b200_prof_timing() { return g_prof_gemm; }      // per-launch event timing is on (b200_prof_begin(1) .. b200_prof_end)
struct GemmEv { cudaEvent_t a, b; double flops, bytes; int kind; };
constexpr int kProfKinds = 8;
static double g_kind_ms[kProfKinds], g_kind_flops[kProfKinds], g_kind_bytes[kProfKinds];
static long long g_kind_launches[kProfKinds];
static std::vector<GemmEv> g_gemm_events;
static std::vector<cudaEvent_t> g_event_pool;
static double g_last_gemm_bytes = 0.0;

void b200_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static cudaEvent_t take_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
bool b200_prof_kind_begin(cudaStream_t stream, int kind, double flops, double bytes) {
  if (!g_prof_gemm || kind < 0 || kind >= kProfKinds) return false;
  GemmEv ev{take_event(), take_event(), flops, bytes, kind};
  cudaEventRecord(ev.a, stream);
  g_gemm_events.push_back(ev);
  return true;
}
void b200_prof_kind_end(cudaStream_t stream) { cudaEventRecord(g_gemm_events.back().b, stream); }

extern "C" int b200_prof_begin(int time_gemm_launches) {
  g_launches.store(0);
  for (auto& e : g_gemm_events) { g_event_pool.push_back(e.a); g_event_pool.push_back(e.b); }
  g_gemm_events.clear();
  g_prof_gemm = time_gemm_launches != 0;
  return B200_OK;
}
// Synchronises the device.  gemm_ms / gemm_flops / gemm_launches describe the timed tcgen05 GEMM launches since
// b200_prof_begin (zeros if timing was off); total_launches counts every kernel launch of the library.
extern "C" int b200_prof_end(double* gemm_ms, double* gemm_flops, long long* gemm_launches, long long* total_launches) {
  g_prof_gemm = false;
  B200_CHECK_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < kProfKinds; ++k) { g_kind_ms[k] = g_kind_flops[k] = g_kind_bytes[k] = 0.0; g_kind_launches[k] = 0; }
  for (auto& e : g_gemm_events) {
    float t = 0.f;
    B200_CHECK_CUDA(cudaEventElapsedTime(&t, e.a, e.b));
    g_kind_ms[e.kind] += t; g_kind_flops[e.kind] += e.flops; g_kind_bytes[e.kind] += e.bytes; ++g_kind_launches[e.kind];
  }
  if (gemm_ms) *gemm_ms = g_kind_ms[0];
  if (gemm_flops) *gemm_flops = g_kind_flops[0];
  g_last_gemm_bytes = g_kind_bytes[0];
  if (gemm_launches) *gemm_launches = g_kind_launches[0];
  if (total_launches) *total_launches = g_launches.load();
  return B200_OK;
}

// Algorithmic HBM bytes (operands + outputs, each once) of the GEMM launches timed by the last b200_prof_begin(1) ..
// b200_prof_end pair; call it after b200_prof_end.
extern "C" int b200_prof_gemm_bytes(double* bytes) {
  if (bytes) *bytes = g_last_gemm_bytes;
  return B200_OK;
}

// Per-kind totals of the launches timed by the last b200_prof_begin(1) .. b200_prof_end pair (call after b200_prof_end):
// arrays of n_kinds entries indexed by B200_PROF_*.
extern "C" int b200_prof_kernels(int n_kinds, double* ms, double* flops, double* bytes, long long* launches) {
  B200_REQUIRE(n_kinds >= 1 && n_kinds <= kProfKinds, "prof_kernels: n_kinds must be in [1, %d]", kProfKinds);
  for (int k = 0; k < n_kinds; ++k) {
    if (ms) ms[k] = g_kind_ms[k];
    if (flops) flops[k] = g_kind_flops[k];
    if (bytes) bytes[k] = g_kind_bytes[k];
    if (launches) launches[k] = g_kind_launches[k];
  }
  return B200_OK;
}

extern "C" int b200_device_check(void) {
  int dev = 0;
  B200_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  B200_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B200_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) return b200_set_error(B200_ERR_ARCH, "device %d is sm_%d%d; this library is sm_100a only", dev, major, minor);
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// TMA descriptor encoding through the driver entry point (no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
namespace gemm {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_2d(CUtensorMap* map, bool is_bf16, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                   uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return b200_set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return b200_set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu pitch=%llu box=%ux%u",
                          (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems,
                          box_inner, box_outer);
  return B200_OK;
}

int encode_tmap_out(CUtensorMap* map, int elem_bytes, const void* ptr, uint64_t cols, uint64_t rows, uint64_t splits,
                    uint64_t pitch_elems, uint64_t split_stride_elems, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return b200_set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {cols, rows, splits};
  cuuint64_t strides[2] = {pitch_elems * elem_bytes, split_stride_elems * elem_bytes};
  cuuint32_t box[3] = {box_cols, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const uint32_t row_bytes = box_cols * elem_bytes;
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  if (row_bytes != 128 && row_bytes != 64 && row_bytes != 32)
    return b200_set_error(B200_ERR_INVALID, "output box row of %u B unsupported", row_bytes);
  CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : (elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32), 3, const_cast<void*>(ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return b200_set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled(out) failed (%d) cols=%llu rows=%llu splits=%llu pitch=%llu box=%u", (int)r,
                          (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)splits, (unsigned long long)pitch_elems, box_cols);
  return B200_OK;
}

// Tile width.  The epilogue stores 32-row boxes whose rows are at most 128 B; 128-B rows (64 bf16 columns) are what the TMA
// store path and L2 like, narrower rows roughly halve the achievable write bandwidth.  So widths are multiples of 64,
// padded if necessary (fully out-of-range boxes are skipped, partially out-of-range ones clipped by the TMA).  For
// small-K, store-bound layers an even number of boxes per tile keeps the two epilogue warps of a lane quarter balanced;
// for large-K layers the widest tile wins (MMA efficiency, fewer re-reads of A).
int pick_block_n(int N, int K) {
  if (N <= 64) return (N + 15) / 16 * 16;
  if (N <= kMaxBlockN) return (N % 32 == 0) ? N : (N + 63) / 64 * 64;   // exact width when the boxes allow it (measured, r01n: N = 96
                                                                         // as one 96-wide tile beats the padded 128: -4 .. -22 %)
  const int order_store[4] = {256, 128, 192, 64};
  const int order_math[4] = {256, 192, 128, 64};
  const int* order = (K <= 192) ? order_store : order_math;
  for (int i = 0; i < 4; ++i)
    if (N % order[i] == 0) return order[i];
  int best = kMaxBlockN;
  long long best_pad = -1;
  for (int d = kMaxBlockN; d >= 128; d -= 64) {
    long long pad = 1LL * ((N + d - 1) / d) * d - N;
    if (best_pad < 0 || pad < best_pad) { best_pad = pad; best = d; }
  }
  return best;
}

// ---------------------------------------------------------------------------------------------
// epilogue for linear layers.  Each epilogue thread owns one output row and gets 16 consecutive
// fp32 accumulator columns per call.
// ---------------------------------------------------------------------------------------------
template <int MODE>
struct EpiLinear {
  struct Params {
    const float* bias;                   // [N] or null
  };

  // v: 16 consecutive accumulator columns of one output row; ax: the matching 16 values of the aux input (RESID: residual,
  // DGELU: the GELU derivative saved by the forward) -> o (and o2 = gelu'(pre-activation) in GELU mode: the backward
  // epilogue is then one multiply instead of a second erf evaluation).  Rows >= M / columns >= N are computed on
  // zero-filled inputs and clipped by the TMA store.
  static bool wants_columns(const Params& ep) { return ep.bias != nullptr; }
  // smem copy of the bias, zero-padded to the tile grid
  __device__ static __forceinline__ void stage_columns(const Params& ep, const CoreParams& p, float* dst, int tid, int nthreads) {
    const int n_pad = p.n_blocks * p.block_n;
    for (int c = tid; c < n_pad; c += nthreads) dst[c] = c < p.N ? __ldg(ep.bias + c) : 0.0f;
  }

  template <bool DUAL>
  __device__ static __forceinline__ void compute(const Params& ep, const CoreParams& p, const float* bias_s, int /*row*/, int col,
                                                 const float (&v)[16], const float (&ax)[16], float (&o)[16], float (&o2)[16]) {
    // packed fp32 pairs throughout: the epilogue is bound by instruction issue
    f32x2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk2(v[2 * i], v[2 * i + 1]);
    if (p.bias_smem) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {                                  // broadcast reads: every lane wants the same 16 floats
        const float4 b = *reinterpret_cast<const float4*>(bias_s + col + g * 4);
        a[g * 2 + 0] = add2(a[g * 2 + 0], pk2(b.x, b.y));
        a[g * 2 + 1] = add2(a[g * 2 + 1], pk2(b.z, b.w));
      }
    } else if (ep.bias != nullptr) {
#pragma unroll
      for (int g = 0; g < 2; ++g)
        if (col + g * 8 < p.N) {                                   // N % 8 == 0 is required by the launcher
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + col + g * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col + g * 8 + 4));
          a[g * 4 + 0] = add2(a[g * 4 + 0], pk2(b0.x, b0.y)); a[g * 4 + 1] = add2(a[g * 4 + 1], pk2(b0.z, b0.w));
          a[g * 4 + 2] = add2(a[g * 4 + 2], pk2(b1.x, b1.y)); a[g * 4 + 3] = add2(a[g * 4 + 3], pk2(b1.z, b1.w));
        }
    }
    if (MODE == B200_EPI_GELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f32x2 y, dy;
        gelu_erf_both2(a[i], y, dy);
        a[i] = y;
        if (DUAL) upk2(dy, o2[2 * i], o2[2 * i + 1]);
      }
    } else if (MODE == B200_EPI_RESID) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = add2(a[i], pk2(ax[2 * i], ax[2 * i + 1]));
    } else if (MODE == B200_EPI_DGELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = mul2(a[i], pk2(ax[2 * i], ax[2 * i + 1]));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) upk2(a[i], o[2 * i], o[2 * i + 1]);
  }
};

// ---------------------------------------------------------------------------------------------
// Convolution + folded BatchNorm epilogue of the eval-mode ResNet path (b200/convnet.py): the BatchNorm scale is folded into
// the bf16 weights, its shift is the bias, and   out = ring(row) ? 0 : [relu](acc + bias (+ aux)).   Every thread of the
// epilogue owns one output row, so the ring test (row -> (y, x) on the padded grid, two fast divisions) costs nothing
// per element; zeroing the ring here is what lets the next 3x3 convolution read its padding from the same tensor.
// ---------------------------------------------------------------------------------------------
template <bool WITH_AUX>
struct EpiConvBN {
  using Lin = EpiLinear<WITH_AUX ? B200_EPI_RESID : B200_EPI_STORE>;
  struct Params {
    const float* bias;
    int relu;
    int H, W, Wp, P;           // padded grid: Wp = W + 2, P = (H + 2) * Wp; H = 0: no ring
    FastDiv div_p, div_wp;
  };
  static bool wants_columns(const Params& ep) { return ep.bias != nullptr; }
  __device__ static __forceinline__ void stage_columns(const Params& ep, const CoreParams& p, float* dst, int tid, int nthreads) {
    Lin::stage_columns(typename Lin::Params{ep.bias}, p, dst, tid, nthreads);
  }
  template <bool DUAL>
  __device__ static __forceinline__ void compute(const Params& ep, const CoreParams& p, const float* bias_s, int row, int col,
                                                 const float (&v)[16], const float (&ax)[16], float (&o)[16], float (&o2)[16]) {
    Lin::template compute<false>(typename Lin::Params{ep.bias}, p, bias_s, row, col, v, ax, o, o2);
    bool keep = true;
    if (ep.H > 0) {
      const uint32_t r = static_cast<uint32_t>(row);
      const uint32_t q = r - ep.div_p.div(r) * static_cast<uint32_t>(ep.P);
      const uint32_t y = ep.div_wp.div(q), x = q - y * static_cast<uint32_t>(ep.Wp);
      keep = y >= 1u && y <= static_cast<uint32_t>(ep.H) && x >= 1u && x <= static_cast<uint32_t>(ep.W);
    }
    const float lo = ep.relu ? 0.0f : -INFINITY;
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = keep ? fmaxf(o[i], lo) : 0.0f;
  }
};

// ---------------------------------------------------------------------------------------------
// ArcFace / CosFace logits epilogue: operands are the unit-norm bf16 rows of the embeddings and of
// the class weights, so the accumulator IS cos(theta).  Margin on the label column, then * s.
// ---------------------------------------------------------------------------------------------
struct EpiMargin {
  struct Params {
    float* cos_label;                     // [B] cos(theta) at the label column (for backward)
    const long long* label;               // [B] int64
    float s, cos_m, sin_m, th, mm, m;
    int kind;                             // 0 = ArcFace, 1 = CosFace (AddMarginProduct)
    int easy_margin;
  };
  static bool wants_columns(const Params&) { return false; }
  __device__ static __forceinline__ void stage_columns(const Params&, const CoreParams&, float*, int, int) {}
  template <bool DUAL>
  __device__ static __forceinline__ void compute(const Params& ep, const CoreParams& p, const float*, int row, int col,
                                                 const float (&v)[16], const float (&)[16], float (&o)[16], float (&)[16]) {
    const bool live = row < p.M && col < p.N;
    const int lab = live ? static_cast<int>(ep.label[row]) : -1;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float c = v[i];
      if (col + i == lab) {
        ep.cos_label[row] = c;
        if (ep.kind == 1) {
          c = c - ep.m;
        } else {
          // reference takes sqrt(1 - cos^2) unclamped (NaN when rounding gives |cos| > 1); clamped here
          const float sine = sqrtf(fmaxf(1.0f - c * c, 0.0f));
          const float phi = c * ep.cos_m - sine * ep.sin_m;
          c = ep.easy_margin ? (c > 0.0f ? phi : c) : (c > ep.th ? phi : c - ep.mm);
        }
      }
      o[i] = c * ep.s;
    }
  }
};

}  // namespace gemm

// ---------------------------------------------------------------------------------------------
// split-K reduce: out[i] (+)= sum_s partial[s][i], fixed order -> deterministic
// ---------------------------------------------------------------------------------------------
// SL split-lanes cooperate on one float4 column group: lane sx sums splits sx, sx+SL, ... in order, then a fixed
// smem tree folds the SL partial sums.  The association order depends only on (splits, SL): deterministic.
// blockIdx.y selects one of up to three outputs whose partials sit side by side ([splits][ny][n]): LayerNorm's dgamma /
// dbeta / bias-gradient rows are folded by one launch.
struct ReduceOuts { float* p[3]; };

template <int SL>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, const ReduceOuts outs, long long n,
                                                            int splits, int accumulate, long long stride) {
  pdl_grid_sync();
  float* __restrict__ out = outs.p[blockIdx.y];
  partial += blockIdx.y * n;
  constexpr int CG = 256 / SL;                       // column groups (of 4 floats) per block
  __shared__ float4 red[SL][CG];
  const int cx = threadIdx.x % CG, sx = threadIdx.x / CG;
  const long long col = (1LL * blockIdx.x * CG + cx) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < n)
    for (int s = sx; s < splits; s += SL) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(partial + 1LL * s * stride + col));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  if (SL > 1) {
    red[sx][cx] = acc;
    __syncthreads();
#pragma unroll
    for (int o = SL / 2; o > 0; o >>= 1) {
      if (sx < o) {
        const float4 a = red[sx][cx], b = red[sx + o][cx];
        red[sx][cx] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      }
      __syncthreads();
    }
    acc = red[0][cx];
  }
  if (sx == 0 && col < n) {
    if (accumulate) {
      const float4 c = *reinterpret_cast<const float4*>(out + col);
      acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
    }
    *reinterpret_cast<float4*>(out + col) = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Batched reduce: between b200_reduce_defer_begin() and b200_reduce_flush() every fixed-order reduction the library would
// launch (split-K weight gradients, LayerNorm / bias-gradient partial rows, the rel-pos table rows) is only RECORDED; the
// flush folds all of them in ONE launch.  A Swin block's backward has eight such reductions of 4-12 us each, latency-bound
// single-wave launches; batched they are one launch per block that fills the machine.  The caller must give every
// recorded reduction its own partial buffer and keep it alive until the flush.
// Summation order per output element: SL split-lanes (lane s takes splits s, s + SL, ...; SL = 1, 8 or 32 from the job's own
// shape) then a fixed smem tree: deterministic, and independent of what else is in the batch.
// ---------------------------------------------------------------------------------------------
struct ReduceJob {
  const float* partial;
  float* out[3];
  long long n, stride;          // floats per output / between consecutive splits
  int splits, ny, accumulate, block0, vec, sl;   // sl: split-lanes per column group (1, 8 or 32)
};
constexpr int kMaxReduceJobs = 56;
struct ReduceBatch {
  ReduceJob job[kMaxReduceJobs];
  int n_jobs, n_blocks;
};
static thread_local ReduceBatch g_batch;
static thread_local bool g_batch_on = false;

__global__ void __launch_bounds__(256) reduce_batch_kernel(const __grid_constant__ ReduceBatch b) {
  pdl_grid_sync();
  __shared__ float4 red[256];
  int j = 0;
  while (j + 1 < b.n_jobs && static_cast<int>(blockIdx.x) >= b.job[j + 1].block0) ++j;
  const ReduceJob& q = b.job[j];
  const int SL = q.sl, CG = 256 / SL;                 // split-lanes x column groups (of 4 floats) of this CTA
  const int per_y = static_cast<int>((q.n + 4 * CG - 1) / (4 * CG));
  const int rel = static_cast<int>(blockIdx.x) - q.block0;
  const int y = rel / per_y;
  const int cx = threadIdx.x % CG, sx = threadIdx.x / CG;
  const long long col = (1LL * (rel - y * per_y) * CG + cx) * 4;
  const float* __restrict__ partial = q.partial + y * q.n;
  float* __restrict__ out = q.out[y];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q.vec) {
    if (col < q.n)
#pragma unroll 8
      for (int s = sx; s < q.splits; s += SL) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(partial + 1LL * s * q.stride + col));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
  } else {
    for (int s = sx; s < q.splits; s += SL) {
      const float* row = partial + 1LL * s * q.stride + col;
      if (col + 0 < q.n) acc.x += __ldg(row + 0);
      if (col + 1 < q.n) acc.y += __ldg(row + 1);
      if (col + 2 < q.n) acc.z += __ldg(row + 2);
      if (col + 3 < q.n) acc.w += __ldg(row + 3);
    }
  }
  if (SL > 1) {                                        // CTA-uniform: the job is a property of the CTA
    red[sx * CG + cx] = acc;
    __syncthreads();
    for (int o = SL / 2; o > 0; o >>= 1) {
      if (sx < o) {
        const float4 a = red[sx * CG + cx], c = red[(sx + o) * CG + cx];
        red[sx * CG + cx] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
      }
      __syncthreads();
    }
    acc = red[cx];
  }
  if (sx == 0 && col < q.n) {
    if (q.vec) {
      if (q.accumulate) {
        const float4 c = *reinterpret_cast<const float4*>(out + col);
        acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
      }
      *reinterpret_cast<float4*>(out + col) = acc;
    } else {
      const float v[4] = {acc.x, acc.y, acc.z, acc.w};
      for (int i = 0; i < 4; ++i)
        if (col + i < q.n) out[col + i] = q.accumulate ? out[col + i] + v[i] : v[i];
    }
  }
}

extern "C" int b200_reduce_defer_begin(void) {
  g_batch.n_jobs = 0;
  g_batch.n_blocks = 0;
  g_batch_on = true;
  return B200_OK;
}
extern "C" int b200_reduce_pending(void) { return g_batch_on ? g_batch.n_jobs : 0; }
// launches the recorded reductions (one kernel) and leaves deferred mode; keep_deferring != 0 re-enters it right away
extern "C" int b200_reduce_flush(void* stream, int keep_deferring) {
  if (g_batch_on && g_batch.n_jobs > 0) {
    double by = 0.0;
    for (int j = 0; j < g_batch.n_jobs; ++j) by += 4.0 * g_batch.job[j].ny * g_batch.job[j].n * (g_batch.job[j].splits + 1 + (g_batch.job[j].accumulate ? 1 : 0));
    const bool prof = b200_prof_kind_begin(reinterpret_cast<cudaStream_t>(stream), B200_PROF_REDUCE, 0.0, by);
    launch_pdl(reduce_batch_kernel, dim3(static_cast<unsigned>(g_batch.n_blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), g_batch);
    if (prof) b200_prof_kind_end(reinterpret_cast<cudaStream_t>(stream));
    B200_LAUNCH_CHECK();
  }
  g_batch.n_jobs = 0;
  g_batch.n_blocks = 0;
  g_batch_on = keep_deferring != 0;
  return B200_OK;
}

// general (also unaligned / n % 4 != 0) recording entry: the rel-pos table rows are 169 floats
int reduce_or_defer(const float* partial, float* const* outs, int ny, long long n, int splits, int accumulate, cudaStream_t stream,
                    long long stride);

int splitk_reduce_multi(const float* partial, float* const* outs, int ny, long long n, int splits, int accumulate, cudaStream_t stream,
                        long long stride) {
  if (stride <= 0) stride = n;
  if (g_batch_on) return reduce_or_defer(partial, outs, ny, n, splits, accumulate, stream, stride);
  B200_REQUIRE(n % 4 == 0 && stride % 4 == 0 && ny >= 1 && ny <= 3, "splitk_reduce: n %% 4 != 0");
  ReduceOuts ro{{outs[0], ny > 1 ? outs[1] : nullptr, ny > 2 ? outs[2] : nullptr}};
  const long long groups = n / 4;
  const unsigned y = static_cast<unsigned>(ny);
  if (splits <= 4) {
    launch_pdl(splitk_reduce_kernel<1>, dim3((unsigned)((groups + 255) / 256), y), dim3(256), 0, stream, partial, ro, n, splits, accumulate, stride);
  } else if (splits <= 32 || groups >= 65536) {
    launch_pdl(splitk_reduce_kernel<8>, dim3((unsigned)((groups + 31) / 32), y), dim3(256), 0, stream, partial, ro, n, splits, accumulate, stride);
  } else {
    launch_pdl(splitk_reduce_kernel<32>, dim3((unsigned)((groups + 7) / 8), y), dim3(256), 0, stream, partial, ro, n, splits, accumulate, stride);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int splitk_reduce(const float* partial, float* out, long long n, int splits, int accumulate, cudaStream_t stream, long long stride) {
  return splitk_reduce_multi(partial, &out, 1, n, splits, accumulate, stream, stride);
}

int reduce_or_defer(const float* partial, float* const* outs, int ny, long long n, int splits, int accumulate, cudaStream_t stream,
                    long long stride) {
  if (stride <= 0) stride = n;
  B200_REQUIRE(ny >= 1 && ny <= 3 && n > 0 && splits >= 1, "reduce: bad job (ny=%d n=%lld splits=%d)", ny, n, splits);
  const bool deferred = g_batch_on;
  if (deferred && g_batch.n_jobs == kMaxReduceJobs) {          // table full: fold what is recorded, keep recording
    int rc = b200_reduce_flush(stream, 1);
    if (rc) return rc;
  }
  if (!deferred) {
    g_batch.n_jobs = 0;
    g_batch.n_blocks = 0;
  }
  ReduceJob& q = g_batch.job[g_batch.n_jobs];
  q.partial = partial;
  bool aligned = (reinterpret_cast<uintptr_t>(partial) & 15) == 0 && n % 4 == 0 && stride % 4 == 0;
  for (int y = 0; y < 3; ++y) {
    q.out[y] = y < ny ? outs[y] : nullptr;
    if (y < ny) aligned = aligned && (reinterpret_cast<uintptr_t>(outs[y]) & 15) == 0;
  }
  q.n = n; q.stride = stride; q.splits = splits; q.ny = ny; q.accumulate = accumulate; q.vec = aligned ? 1 : 0;
  // Split-lanes per column group, from the job's own shape: few splits (the wide weight gradients of stages 3-4) -> one
  // thread sums all splits of its float4 (independent loads, no smem, 1024 columns per CTA); hundreds of partial rows of a
  // narrow output (LayerNorm / bias rows) -> 32 lanes keep the serial chain short.
  q.sl = splits <= 16 ? 1 : (splits <= 64 ? 8 : 32);
  q.block0 = g_batch.n_blocks;
  g_batch.n_blocks += ny * static_cast<int>((n + 4 * (256 / q.sl) - 1) / (4 * (256 / q.sl)));
  ++g_batch.n_jobs;
  if (!deferred) {                                              // immediate mode: a batch of one
    g_batch_on = true;
    return b200_reduce_flush(stream, 0);
  }
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
static int gemm_tn_impl(const void* a, long long lda, const void* b, long long ldb, int M, int N, int K, int is_bf16,
                        int mode, void* out, long long ldo, int out_fp32, void* out2, long long ldo2, const float* bias,
                        const void* aux, long long ldaux, int splits, long long split_stride, int block_n, void* stream,
                        int taps, const int* tap_shift) {
  B200_REQUIRE(N % 8 == 0, "gemm_tn: N must be a multiple of 8 (got %d)", N);
  B200_REQUIRE(mode >= B200_EPI_STORE && mode <= B200_EPI_DGELU_Q8, "gemm_tn: bad epilogue mode %d", mode);
  const bool q8 = mode == B200_EPI_GELU_Q8 || mode == B200_EPI_DGELU_Q8;      // the GELU derivative travels as 8-bit codes
  if (mode == B200_EPI_GELU_Q8) { mode = B200_EPI_GELU; B200_REQUIRE(out2 != nullptr, "gemm_tn: GELU_Q8 needs out2"); }
  if (mode == B200_EPI_DGELU_Q8) mode = B200_EPI_DGELU;
  if (mode == B200_EPI_RESID || mode == B200_EPI_DGELU) B200_REQUIRE(aux != nullptr && ldaux % (q8 ? 16 : 8) == 0, "gemm_tn: mode needs aux");
  if (mode != B200_EPI_PARTIAL) B200_REQUIRE(splits <= 1, "gemm_tn: split-K only with the PARTIAL epilogue");
  if (mode != B200_EPI_GELU) out2 = nullptr;
  const int eb = (out_fp32 || mode == B200_EPI_PARTIAL) ? 4 : 2;
  gemm::Operands o{a, (int)lda, b, (int)ldb, M, N, K, is_bf16 != 0, block_n, splits, 0};
  o.taps = taps;
  for (int t = 0; t < taps; ++t) o.tap_shift[t] = tap_shift[t];
  const bool use_aux = mode == B200_EPI_RESID || mode == B200_EPI_DGELU;
  B200_REQUIRE(!(use_aux && eb != 2), "gemm_tn: RESID / DGELU epilogues write bf16");
  B200_REQUIRE(!(mode == B200_EPI_GELU && eb != 2), "gemm_tn: the GELU epilogue writes bf16");
  gemm::Output od{out, ldo, eb, out2, ldo2, mode == B200_EPI_PARTIAL ? split_stride : 0, use_aux ? aux : nullptr, use_aux ? ldaux : 0};
  auto st = reinterpret_cast<cudaStream_t>(stream);
  // CTA pairs for the compute-bound layers: K-major, unsplit, K >= 384 (six or more K blocks), whole tiles of >= 192 columns,
  // at least two waves of row-block pairs (stages 3-4 at the bench's batch).  See gemm_core.cuh: CG2.
  {
    const int bn = block_n > 0 ? block_n : gemm::pick_block_n(N, K);
    static const int min_k = [] { const char* e = getenv("B200_CG2_MINK"); return e != nullptr ? atoi(e) : 384; }();      // experiments
    const bool pair_ok = gemm::b200_cg2() && is_bf16 && K >= min_k && bn >= 192 && bn % 32 == 0 && N % bn == 0 && M >= 2 * 128 && eb == 2 && !q8;
    if (pair_ok) {
      switch (mode) {
        case B200_EPI_STORE: return gemm::launch<gemm::EpiLinear<B200_EPI_STORE>, 2, false, false, false, 8, true>(o, od, {bias}, st);
        case B200_EPI_GELU:
          if (out2 != nullptr) return gemm::launch<gemm::EpiLinear<B200_EPI_GELU>, 2, true, false, false, 8, true>(o, od, {bias}, st);
          return gemm::launch<gemm::EpiLinear<B200_EPI_GELU>, 2, false, false, false, 8, true>(o, od, {bias}, st);
        case B200_EPI_RESID: return gemm::launch<gemm::EpiLinear<B200_EPI_RESID>, 2, false, true, false, 8, true>(o, od, {bias}, st);
        case B200_EPI_DGELU: return gemm::launch<gemm::EpiLinear<B200_EPI_DGELU>, 2, false, true, false, 8, true>(o, od, {bias}, st);
        default: break;
      }
    }
  }
  switch (mode) {
    case B200_EPI_STORE:
      // small-K bf16 layers (K <= 4 blocks of 64): sixteen epilogue warps (B200_EPI16=0 -> eight).  Measured (r01y): qkv of stage 1
      // 143.5 -> 134.8 us, out-projection data gradient 60.1 -> 57.7 us, nothing elsewhere: the warp count is not what starves these layers
      if (eb == 2 && K <= 256 && b200_epi16()) return gemm::launch<gemm::EpiLinear<B200_EPI_STORE>, 2, false, false, false, 16>(o, od, {bias}, st);
      if (eb == 2) return gemm::launch<gemm::EpiLinear<B200_EPI_STORE>, 2, false, false>(o, od, {bias}, st);
      return gemm::launch<gemm::EpiLinear<B200_EPI_STORE>, 4, false, false>(o, od, {bias}, st);
    case B200_EPI_GELU:
      if (out2 != nullptr && q8) return gemm::launch<gemm::EpiLinear<B200_EPI_GELU>, 2, true, false, true>(o, od, {bias}, st);
      if (out2 != nullptr) return gemm::launch<gemm::EpiLinear<B200_EPI_GELU>, 2, true, false>(o, od, {bias}, st);
      return gemm::launch<gemm::EpiLinear<B200_EPI_GELU>, 2, false, false>(o, od, {bias}, st);
    case B200_EPI_RESID: return gemm::launch<gemm::EpiLinear<B200_EPI_RESID>, 2, false, true>(o, od, {bias}, st);
    case B200_EPI_DGELU:
      if (q8) return gemm::launch<gemm::EpiLinear<B200_EPI_DGELU>, 2, false, true, true>(o, od, {bias}, st);
      return gemm::launch<gemm::EpiLinear<B200_EPI_DGELU>, 2, false, true>(o, od, {bias}, st);
    default: return gemm::launch<gemm::EpiLinear<B200_EPI_PARTIAL>, 4, false, false>(o, od, {bias}, st);
  }
}

extern "C" int b200_gemm_tn(const void* a, long long lda, const void* b, long long ldb, int M, int N, int K, int is_bf16,
                            int mode, void* out, long long ldo, int out_fp32, void* out2, long long ldo2, const float* bias,
                            const void* aux, long long ldaux, int splits, long long split_stride, int block_n, void* stream) {
  return gemm_tn_impl(a, lda, b, ldb, M, N, K, is_bf16, mode, out, ldo, out_fp32, out2, ldo2, bias, aux, ldaux, splits, split_stride, block_n,
                      stream, 0, nullptr);
}

// Implicit convolution on the tensor cores: out[r, n] = sum_t sum_c a[r + tap_shift[t], c] * b[n, t * C + c] (+ aux[r, n] with
// B200_EPI_RESID) for bf16 a [M, C] (C a multiple of 64) and b [N, taps * C].  With the activations of a padded NHWC grid stored
// as rows ((image, y, x) -> row, one zero ring around every image) a 3x3 convolution is the nine shifts (dy * padded_width + dx):
// no im2col matrix exists anywhere, the TMA producer reads the shifted row blocks straight from the activation tensor (rows
// outside [0, M) come back as zeros).  torchvision.models.resnet50 conv2 of every Bottleneck, as configs/dog_fe/fe_dogs_config.py:96-109 uses it.
extern "C" int b200_gemm_taps(const void* a, long long lda, const void* b, long long ldb, int M, int N, int C, int taps, const int* tap_shift,
                              int mode, void* out, long long ldo, const void* aux, long long ldaux, void* stream) {
  B200_REQUIRE(taps >= 1 && taps <= 9 && tap_shift != nullptr && C % 64 == 0, "gemm_taps: taps in [1, 9], C a multiple of 64 (taps=%d C=%d)", taps, C);
  B200_REQUIRE(mode == B200_EPI_STORE || mode == B200_EPI_RESID, "gemm_taps: STORE or RESID epilogue");
  return gemm_tn_impl(a, lda, b, ldb, M, N, taps * C, 1, mode, out, ldo, 0, nullptr, 0, nullptr, aux, ldaux, 1, 0, 0, stream, taps, tap_shift);
}

// Eval-mode convolution + BatchNorm (+ residual) (+ ReLU) in one launch: out[r, n] = ring(r) ? 0 : act(sum a . b + bias[n] (+ aux[r, n])),
// a [M, C] rows of the padded (H, W) grid, b [N, taps * C] = the convolution weights with the BatchNorm scale folded in,
// bias = the BatchNorm shift.  taps = 0 / 1: a 1x1 convolution (plain GEMM, K = C); taps = 9 with tap_shift: 3x3 (b200_gemm_taps).
extern "C" int b200_gemm_conv_bn(const void* a, long long lda, const void* b, long long ldb, int M, int N, int C, int taps, const int* tap_shift,
                                 const float* bias, int relu, int H, int W, const void* aux, long long ldaux, void* out, long long ldo,
                                 void* stream) {
  B200_REQUIRE(N % 8 == 0 && C % 8 == 0 && taps >= 0 && taps <= 9, "gemm_conv_bn: bad shape N=%d C=%d taps=%d", N, C, taps);
  if (taps <= 1) taps = 0;
  B200_REQUIRE(taps == 0 || (tap_shift != nullptr && C % 64 == 0), "gemm_conv_bn: a tap convolution needs tap_shift and C %% 64 == 0");
  B200_REQUIRE(aux == nullptr || ldaux % 8 == 0, "gemm_conv_bn: aux pitch must be a multiple of 8 elements");
  const int K = taps > 0 ? taps * C : C;
  gemm::Operands o{a, (int)lda, b, (int)ldb, M, N, K, true, 0, 1, 0};
  o.taps = taps;
  for (int t = 0; t < taps; ++t) o.tap_shift[t] = tap_shift[t];
  gemm::Output od{out, ldo, 2, nullptr, 0, 0, aux, aux != nullptr ? ldaux : 0};
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const int Wp = W + 2, P = (H + 2) * (W + 2);
  const int bn = gemm::pick_block_n(N, K);
  const bool pair = gemm::b200_cg2() && K >= 384 && bn >= 192 && bn % 32 == 0 && N % bn == 0 && M >= 2 * 128;      // as b200_gemm_tn
  if (aux != nullptr) {
    gemm::EpiConvBN<true>::Params ep{bias, relu, H, W, Wp, P, make_fastdiv(static_cast<uint32_t>(P)), make_fastdiv(static_cast<uint32_t>(Wp))};
    if (pair) return gemm::launch<gemm::EpiConvBN<true>, 2, false, true, false, 8, true>(o, od, ep, st);
    return gemm::launch<gemm::EpiConvBN<true>, 2, false, true>(o, od, ep, st);
  }
  gemm::EpiConvBN<false>::Params ep{bias, relu, H, W, Wp, P, make_fastdiv(static_cast<uint32_t>(P)), make_fastdiv(static_cast<uint32_t>(Wp))};
  if (pair) return gemm::launch<gemm::EpiConvBN<false>, 2, false, false, false, 8, true>(o, od, ep, st);
  return gemm::launch<gemm::EpiConvBN<false>, 2, false, false>(o, od, ep, st);
}

// dW[N,K] (fp32 split partials) = dY[tokens,N]^T * X[tokens,K]: both operands are read in place, MN-major - no transposes.
extern "C" int b200_gemm_wgrad(const void* dy, long long ldy, const void* x, long long ldx, long long tokens, int N, int K,
                               float* partial, int splits, int block_n, void* stream) {
  B200_REQUIRE(K % 8 == 0 && N > 0 && tokens > 0 && tokens < (1LL << 31), "gemm_wgrad: bad shape tokens=%lld N=%d K=%d", tokens, N, K);
  gemm::Operands o{dy, (int)ldy, x, (int)ldx, N, K, static_cast<int>(tokens), true, block_n, splits, 0, true};
  gemm::Output od{partial, K, 4, nullptr, 0, 1LL * N * K, nullptr, 0};
  return gemm::launch<gemm::EpiLinear<B200_EPI_PARTIAL>, 4, false, false>(o, od, {nullptr}, reinterpret_cast<cudaStream_t>(stream));
}

// Weight gradient of an implicit convolution (b200_gemm_taps): partial [splits][N][taps * C] with
// dW[n, t * C + c] = sum_p dy[p, n] * x[p + tap_shift[t], c] (rows outside [0, tokens) read as zero).  One launch for all taps:
// the column blocks of one split walk the same rows at the same time, so the nine shifted reads of x (and the nine of dy) are
// served by L2.
extern "C" int b200_gemm_wgrad_taps(const void* dy, long long ldy, const void* x, long long ldx, long long tokens, int N, int C, int taps,
                                    const int* tap_shift, float* partial, int splits, void* stream) {
  B200_REQUIRE(C % 64 == 0 && N > 0 && tokens > 0 && tokens < (1LL << 31) && taps >= 1 && taps <= 9, "gemm_wgrad_taps: bad shape tokens=%lld N=%d C=%d", tokens, N, C);
  const int bn = 128;        // a tile spans two taps of a 64-channel layer (they share the dy tile); 128 keeps the operand ring deep
  gemm::Operands o{dy, (int)ldy, x, (int)ldx, N, taps * C, static_cast<int>(tokens), true, bn, splits, 0, true};
  o.taps = taps;
  for (int t = 0; t < taps; ++t) o.tap_shift[t] = tap_shift[t];
  gemm::Output od{partial, 1LL * taps * C, 4, nullptr, 0, 1LL * N * taps * C, nullptr, 0};
  return gemm::launch<gemm::EpiLinear<B200_EPI_PARTIAL>, 4, false, false>(o, od, {nullptr}, reinterpret_cast<cudaStream_t>(stream));
}

// The same launch also producing the bias gradient db[N] = colsum(dY) as [splits][N] fp32 partial rows in `colsum_partial`
// (an all-ones MMA inside the kernel: no second pass over dY).  *fused = 0 (and colsum_partial untouched) when the tile
// shape leaves no spare TMEM columns for it (K > 240 with a 256-wide tile): the caller then runs b200_colsum.
extern "C" int b200_gemm_wgrad_bias(const void* dy, long long ldy, const void* x, long long ldx, long long tokens, int N, int K,
                                    float* partial, float* colsum_partial, int splits, int block_n, int* fused, void* stream) {
  B200_REQUIRE(K % 8 == 0 && N > 0 && tokens > 0 && tokens < (1LL << 31), "gemm_wgrad: bad shape tokens=%lld N=%d K=%d", tokens, N, K);
  const int bn = block_n > 0 ? block_n : gemm::pick_block_n(K, static_cast<int>(tokens));
  const bool can = colsum_partial != nullptr && bn <= gemm::kOnesCol;
  if (fused) *fused = can ? 1 : 0;
  gemm::Operands o{dy, (int)ldy, x, (int)ldx, N, K, static_cast<int>(tokens), true, block_n, splits, 0, true};
  gemm::Output od{partial, K, 4, nullptr, 0, 1LL * N * K, nullptr, 0};
  if (can) od.colsum = colsum_partial;
  return gemm::launch<gemm::EpiLinear<B200_EPI_PARTIAL>, 4, false, false>(o, od, {nullptr}, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b200_gemm_splits(int K, int splits) { return gemm::effective_splits(K, splits); }

extern "C" int b200_splitk_reduce(const float* partial, float* out, long long n, int splits, int accumulate, void* stream) {
  return splitk_reduce(partial, out, n, splits, accumulate, reinterpret_cast<cudaStream_t>(stream), n);
}

extern "C" int b200_margin_logits(const void* emb_unit, const void* w_unit, int B, int C, int E, const long long* label,
                                  float s, float m, int kind, int easy_margin, float* logits, long long ldo,
                                  float* cos_label, void* stream) {
  B200_REQUIRE(E % 8 == 0, "margin_logits: embedding size must be a multiple of 8");
  gemm::Operands o{emb_unit, E, w_unit, E, B, C, E, true, 0, 1, 0};
  const double md = static_cast<double>(m), pi = 3.14159265358979323846;   // constants as in large_margin.py:64-67
  B200_REQUIRE(ldo % 4 == 0, "margin_logits: logits pitch must be a multiple of 4");
  gemm::Output od{logits, ldo, 4, nullptr, 0, 0, nullptr, 0};
  gemm::EpiMargin::Params ep{cos_label, label, s, (float)cos(md), (float)sin(md), (float)cos(pi - md),
                             (float)(sin(pi - md) * md), m, kind, easy_margin};
  return gemm::launch<gemm::EpiMargin, 4, false, false>(o, od, ep, reinterpret_cast<cudaStream_t>(stream));
}
