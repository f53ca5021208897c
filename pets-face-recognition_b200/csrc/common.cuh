// Shared device/host helpers for the sm_100a kernels: error plumbing, bf16 packing, warp
// reductions, and thin inline-PTX wrappers for mbarrier / TMA / tcgen05 (no CUTLASS).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <utility>

// ---------------------------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns 0 or a negative code; text via b200_last_error()
// ---------------------------------------------------------------------------------------------
enum B200Status : int {
  B200_OK = 0,
  B200_ERR_INVALID = -1,    // bad argument / unsupported shape
  B200_ERR_CUDA = -2,       // a CUDA runtime/driver call failed
  B200_ERR_WORKSPACE = -3,  // workspace too small
  B200_ERR_ARCH = -4,       // not an sm_100 device
};

int b200_set_error(int code, const char* fmt, ...);

#define B200_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      (void)cudaGetLastError(); /* reported here: must not resurface at the next launch check */ \
      return b200_set_error(B200_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,   \
                            cudaGetErrorString(_e));                                       \
    }                                                                                      \
  } while (0)

#define B200_REQUIRE(cond, ...)                                              \
  do {                                                                       \
    if (!(cond)) return b200_set_error(B200_ERR_INVALID, __VA_ARGS__);       \
  } while (0)

// every kernel launch of the library goes through this: counts launches (b200_prof_end reports them)
void b200_count_launch();
#define B200_LAUNCH_CHECK()                    \
  do {                                         \
    b200_count_launch();                       \
    B200_CHECK_CUDA(cudaGetLastError());       \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel of the library is launched with the stream-serialisation attribute and
// starts with pdl_grid_sync() - griddepcontrol.wait (block until the previous kernel in the stream has completed and its
// memory is visible) followed by griddepcontrol.launch_dependents (let the next kernel's CTAs be scheduled as SMs drain).
// Placed after a kernel's data-independent prologue (barrier init, TMEM allocation, smem tables), this hides launch
// latency and the ramp-down / ramp-up bubbles between the ~500 kernels of a training step.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}
#endif

int b200_num_sms();   // cached cudaDevAttrMultiProcessorCount of the current device (148 on B200)
// out[i] (+)= sum_s partial[s * stride + i], i < n, fixed order (deterministic); stride <= 0 means n
// one launch for a table of fp32 [R, Cc] -> bf16 (+ transposed bf16) casts; offsets are elements from in_base / out_base
struct CastJob { long long in_off, dst_off, dst_t_off; int R, Cc, tile0, tiles_c; };
struct CastJobs { CastJob job[64]; int n; };
int cast_transpose_multi(const float* in_base, void* out_base, const CastJobs& jobs, int total_tiles, cudaStream_t stream);
int splitk_reduce(const float* partial, float* out, long long n, int splits, int accumulate, cudaStream_t stream, long long stride);
// the same reduction without alignment requirements; recorded instead of launched while a reduce batch is open (gemm.cu)
int reduce_or_defer(const float* partial, float* const* outs, int ny, long long n, int splits, int accumulate, cudaStream_t stream,
                    long long stride);
// up to three outputs whose partial rows are adjacent ([splits][ny][n], row pitch `stride`) in one launch
int splitk_reduce_multi(const float* partial, float* const* outs, int ny, long long n, int splits, int accumulate, cudaStream_t stream,
                        long long stride);
// optional per-launch CUDA-event timing of the library's kernels by kind (bench.py roofline): no-ops unless enabled.
// kinds: see B200_PROF_* in b200_fe.h; flops / bytes are the ALGORITHMIC work of the launch
bool b200_prof_kind_begin(cudaStream_t stream, int kind, double flops, double bytes);
bool b200_prof_timing();
void b200_prof_kind_end(cudaStream_t stream);
inline bool b200_prof_gemm_begin(cudaStream_t stream, double flops, double bytes) { return b200_prof_kind_begin(stream, 0, flops, bytes); }
inline void b200_prof_gemm_end(cudaStream_t stream) { b200_prof_kind_end(stream); }

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  bf162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t v) {
  bf162 t = *reinterpret_cast<bf162*>(&v);
  return __bfloat1622float2(t);
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Exact-erf GELU (nn.GELU() default, models/swin.py:41) from the Abramowitz-Stegun 7.1.26 rational approximation of erfc
// (|error| <= 1.5e-7, far below bf16 resolution), arranged for the GEMM epilogues, which are ALU-bound:
//   w(x) = 0.5 erfc(|x| / sqrt 2) = t (a1' + t (a2' + ..)) exp(-x^2 / 2),  t = 1 / (1 + p |x| / sqrt 2),  ai' = ai / 2
//   gelu(x)  = x Phi(x)          = max(x, 0) - |x| w(x)                                   (12 instructions, 2 MUFU)
//   gelu'(x) = Phi(x) + x phi(x) = [x >= 0] - copysign(w, x) + x exp(-x^2 / 2) / sqrt(2 pi)
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void gelu_parts(float x, float& w, float& gauss) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.0f));
  gauss = ex2_approx(-0.72134752044448170f * (x * x));          // exp(-x^2 / 2)
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  w = poly * t * gauss;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float w, g;
  gelu_parts(x, w, g);
  return fmaf(-fabsf(x), w, fmaxf(x, 0.0f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float w, g;
  gelu_parts(x, w, g);
  const float phi_cdf = (x >= 0.0f ? 1.0f : 0.0f) - copysignf(w, x);
  return fmaf(x * 0.39894228040143268f, g, phi_cdf);
}

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: one instruction for two lanes of a 64-bit register pair).  The GEMM
// epilogues are bound by instruction issue, not by the FP32 pipe, so halving the instruction count of the element-wise
// math is what buys time there.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 dup2(float v) { return pk2(v, v); }

// GELU and its derivative for two pre-activations at once (the forward GELU epilogue saves the derivative for the backward
// pass instead of the pre-activation).  Same rational erfc as gelu_parts, rearranged so that everything but the two MUFU
// pairs and the sign handling is a packed instruction:
//   g' = exp(-x^2 / 2) / sqrt(2 pi) = 2^(x * (-k x) + log2(1 / sqrt(2 pi)))        w = t poly'(t) g'   (ai' = ai sqrt(2 pi) / 2)
//   Phi = 1/2 + copysign(1/2 - w, x)        gelu = x Phi        gelu' = Phi + x g'
// 12 instructions per element for both outputs (the scalar form above needs 21).
__device__ __forceinline__ void gelu_erf_both2(f32x2 x, f32x2& y, f32x2& dy) {
  constexpr float kS = 1.2533141373155001f;                                  // sqrt(2 pi) / 2
  float x0, x1;
  upk2(x, x0, x1);
  const f32x2 targ = fma2(dup2(0.3275911f * 0.70710678118654752f), pk2(fabsf(x0), fabsf(x1)), dup2(1.0f));
  float a0, a1;
  upk2(targ, a0, a1);
  const f32x2 t = pk2(rcp_approx(a0), rcp_approx(a1));
  const f32x2 e = fma2(x, mul2(x, dup2(-0.72134752044448170f)), dup2(-1.3257480647361593f));   // log2(1 / sqrt(2 pi))
  float e0, e1;
  upk2(e, e0, e1);
  const f32x2 g = pk2(ex2_approx(e0), ex2_approx(e1));
  f32x2 poly = fma2(dup2(kS * 1.061405429f), t, dup2(kS * -1.453152027f));
  poly = fma2(poly, t, dup2(kS * 1.421413741f));
  poly = fma2(poly, t, dup2(kS * -0.284496736f));
  poly = fma2(poly, t, dup2(kS * 0.254829592f));
  const f32x2 w = mul2(mul2(poly, t), g);
  const f32x2 s = fma2(w, dup2(-1.0f), dup2(0.5f));
  float s0, s1;
  upk2(s, s0, s1);
  const f32x2 phi = add2(pk2(copysignf(s0, x0), copysignf(s1, x1)), dup2(0.5f));
  y = mul2(x, phi);
  dy = fma2(x, g, phi);
}

// Division by a run-time constant as multiply-high + shift (the per-tile / per-task index arithmetic sits on the critical
// path of the small-K GEMMs and of the window-attention kernels: ncu showed the emulated divisions at 17 % of the
// out-projection's and 12 % of the attention forward's stall samples).  Exact for 0 <= n < 2^31.
struct FastDiv {
  uint32_t d, mul, shift;
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1u ? n : (__umulhi(n, mul) >> shift); }
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{d, 0u, 0u};
  if (d <= 1u) { f.d = 1u; return f; }
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;                    // ceil(log2 d) >= 1
  const uint32_t p = 31u + l;
  f.mul = static_cast<uint32_t>(((1ull << p) + d - 1) / d);
  f.shift = p - 32u;
  return f;
}

// Raster row <-> window-major row of a [B, H, W] token grid cut into 7 x 7 windows, optionally after the cyclic shift by 3 of
// the shifted Swin blocks (models/swin.py:8-14, :112): token (r, c) of window (wy, wx) of image b is pixel
// ((7 wy + r + off) mod H, (7 wx + c + off) mod W) and gets window-major row ((b nwh + wy) nww + wx) 49 + 7 r + c.  The
// LayerNorm in front of to_qkv writes its output in this order, so that every (window, head) operand of the attention
// kernels is a contiguous [49 rows x 32 columns] box a TMA descriptor can fetch.
struct WinMap {
  int enabled, H, W, nwh, nww, off;
  FastDiv div_w, div_hw;
};
inline WinMap make_winmap(int H, int W, int shifted) {
  WinMap m;
  m.enabled = 1; m.H = H; m.W = W; m.nwh = H / 7; m.nww = W / 7; m.off = shifted ? 3 : 0;
  m.div_w = make_fastdiv(static_cast<uint32_t>(W));
  m.div_hw = make_fastdiv(static_cast<uint32_t>(H) * static_cast<uint32_t>(W));
  return m;
}
inline WinMap no_winmap() { WinMap m{}; m.enabled = 0; m.div_w = make_fastdiv(1); m.div_hw = make_fastdiv(1); return m; }
#ifdef __CUDACC__
__device__ __forceinline__ long long win_row(const WinMap& m, long long raster_row) {
  const uint32_t row = static_cast<uint32_t>(raster_row);                  // launchers require B * H * W < 2^31
  const uint32_t b = m.div_hw.div(row);
  const uint32_t rem = row - b * static_cast<uint32_t>(m.H * m.W);
  const uint32_t y = m.div_w.div(rem);
  const uint32_t x = rem - y * static_cast<uint32_t>(m.W);
  int ys = static_cast<int>(y) - m.off, xs = static_cast<int>(x) - m.off;
  if (ys < 0) ys += m.H;
  if (xs < 0) xs += m.W;
  const int wy = ys / 7, wx = xs / 7;
  const int r = ys - 7 * wy, c = xs - 7 * wx;
  return (static_cast<long long>(b) * m.nwh + wy) * m.nww * 49 + wx * 49 + r * 7 + c;
}
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// The same wait for the single-thread producer / MMA-issuer roles.  These roles share their warp scheduler with epilogue
// warps whose instruction issue bounds the small-K GEMMs, so a waiting role must not burn issue slots: try_wait with a
// suspend-time hint compiles to SYNCS.TRYWAIT + NANOSLEEP.SYNCS <ns> (a sleep the barrier's completion ends early) instead
// of a probe / branch loop (ncu, r01e: the probe loop with a 32 ns nanosleep was 19 % of all instructions of the fc1 GEMM,
// all of them on two of the four schedulers).  ns == 0 selects the plain probe loop.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, uint32_t ns) {
  if (ns == 0) {
    mbar_wait(bar, parity);
    return;
  }
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITB_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONEB_%=;\n\t"
      "bra WAITB_%=;\n\t"
      "DONEB_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity), "r"(ns)
      : "memory");
}
// suspend-time hint of the single-thread roles (ns); B200_WAIT_NS overrides it for experiments
int b200_wait_ns();
int b200_epi16();   // 1 (default): small-K plain GEMM epilogues run with 16 epilogue warps; B200_EPI16=0 -> 8
// B200_REVERSE=1 (experiment, default off): the streaming kernels between two GEMMs (LayerNorm, window attention) walk their
// rows / tasks in descending order so that producer -> consumer hand-overs of tensors larger than L2 would hit the most
// recently written part.  Measured on B200 (r01m, 2 x 2 runs): no gain (20.1 / 20.4 ms vs 19.9 / 20.1 ms per step).
int b200_reverse_rows();

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) 2-D tile load, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}

// TMA prefetch of a tile into L2 only (no smem, no barrier): decouples HBM latency from the depth of the smem ring
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c_inner),
               "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const void* tmap, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}

// TMA bulk tensor STORE smem -> global (3-D: column, row, split), bulk-group completion
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16/fp16 inputs, fp32 accumulate, one CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster on one TPC share one M = 256 MMA.  Each holds its own 128 rows of A and HALF
// of the B tile; the leader CTA (cluster rank 0) issues the MMAs, both receive their 128 accumulator rows in their own TMEM.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(static_cast<uint16_t>(3))
               : "memory");
}
// the address of the leader CTA's copy of a barrier (the cluster-window address of rank r carries r in bit 24)
__device__ __forceinline__ uint32_t leader_bar(uint32_t bar) { return bar & 0xFEFFFFFFu; }
// TMA load issued by either CTA of the pair whose completion is counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t smem_dst, const void* tmap, uint32_t bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// plain arrive on the copy of a barrier that lives in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(bar),
      "r"(rank)
      : "memory");
}

// TMEM -> registers: this warp's 32 lanes (TMEM lanes 32*(warp%4)..+31), 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {       // one fp32 column
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
// registers -> TMEM: this warp's 32 lanes, 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r) {          // one 32-bit column
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand comes from tensor memory (row i in lane i, two 16-bit K elements per
// 32-bit column, 8 columns per K = 16 step)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major or MN-major operand tile stored with the 128-byte
// swizzle (rows of 128 B, 8-row / 1024-B swizzle atoms, as written by a SWIZZLE_128B TMA box).
//   K-major : canonical ((8,n),2):((8,SBO),1) in 16-B units -> SBO = 1024 B between 8-row groups.
//   MN-major: canonical ((8,n),(8,k)):((1,LBO),(8,SBO))     -> LBO = bytes between 64-element MN
//             chunks, SBO = 1024 B between 8-k-row groups.
// bits: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout=2 (SW128)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
