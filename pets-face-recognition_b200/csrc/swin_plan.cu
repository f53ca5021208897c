// Swin backbone forward / backward as ONE native call each (reference models/swin.py:188-225 and the
// autograd graph PyTorch builds for it).  The plan owns no memory: the caller (PyTorch's caching
// allocator, through b200/plan.py) provides
//   * params   : flat fp32 parameter buffer  (layout = b200_swin_param_offsets, state-dict order, masks excluded)
//   * grads    : flat fp32 gradient buffer, same layout
//   * wcache   : bf16 copies of every GEMM weight, [N,K] and transposed [K,N]   (b200_swin_wcache_bytes)
//   * workspace: activations saved for backward + transients                     (b200_swin_workspace_bytes)
// Activations are bf16, token-major NHWC ([B*H*W, C]) through the whole network - the reference's
// NCHW permute between stages (models/swin.py:193) never happens; statistics, softmax, loss and all
// accumulations are fp32; parameters and gradients are fp32.
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

#include "b200_fe.h"

namespace {

constexpr int kWindow = 7;
constexpr int kHeadDim = 32;

inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

struct ParamRef { long long off = -1; long long numel = 0; };

struct BlockParams {
  ParamRef ln1_w, ln1_b, pos, wqkv, wo, bo, ln2_w, ln2_b, w1, b1, w2, b2;
  long long wqkv16 = 0, wqkv16t = 0, wo16 = 0, wo16t = 0, w116 = 0, w116t = 0, w216 = 0, w216t = 0;   // wcache element offsets
};
struct BlockActs {   // byte offsets into the workspace
  long long mean1, rstd1, xn1, qkv, attn, lse, xmid, mean2, rstd2, xn2, hgrad, hact, xout;
};
struct Stage {
  int C, Hs, heads, df, nblocks, Kp;   // Kp = patch feature count (C_in * df^2)
  long long M;
  ParamRef wp, bp;
  long long wp16 = 0, wp16t = 0;
  long long cols, x0;                  // workspace byte offsets
  std::vector<BlockParams> bp_;
  std::vector<BlockActs> ba_;
};

struct Plan {
  int B, img, channels, hidden, num_classes, training;
  int layers[4], heads[4], df[4];
  Stage st[4];
  ParamRef head_ln_w, head_ln_b, head_w, head_b;
  long long head_w16 = 0, head_w16t = 0;
  long long n_params = 0;              // fp32 elements in the flat buffers
  long long wcache_elems = 0;
  long long ws_bytes = 0;
  std::vector<ParamRef> order;         // state-dict order
  // head / tail activations
  long long pooled, pool_mean, pool_rstd, pooled_n;
  // backward transients
  long long g_a, g_b, d_big, d_small, arena, demb16, dpooled;
  long long arena_bytes;               // partial-sum arena of the batched reductions (one batch per Swin block)
  // backward: the weight-gradient GEMMs are off the critical chain (nothing but the optimizer reads them), so they run on a
  // second stream and fill the tails of the chain's kernels; created on first use (B200_WGRAD_STREAM=0: everything in order)
  mutable cudaStream_t side = nullptr;
  mutable cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_done[4] = {nullptr, nullptr, nullptr, nullptr};
  mutable cudaEvent_t ev_red[2] = {nullptr, nullptr};      // the batched reduction that last read arena half 0 / 1
  mutable int half = 0;                                      // arena half the next partial sums go to (persists across calls)
};

ParamRef add_param(Plan& p, long long numel) {
  ParamRef r;
  r.off = p.n_params;
  r.numel = numel;
  p.n_params = align_up(p.n_params + numel, 64);   // 256-B aligned tensors: float4 / TMA-safe
  p.order.push_back(r);
  return r;
}
long long add_w16(Plan& p, long long numel) {
  long long o = p.wcache_elems;
  p.wcache_elems = align_up(p.wcache_elems + numel, 128);
  return o;
}
long long add_ws(Plan& p, long long bytes) {
  long long o = p.ws_bytes;
  p.ws_bytes = align_up(p.ws_bytes + bytes, 256);
  return o;
}

void build(Plan& p) {
  int c_in = p.channels, res = p.img;
  // ---- parameters, in the registration order of models/swin.py:201-217 (masks are not parameters here)
  for (int s = 0; s < 4; ++s) {
    Stage& S = p.st[s];
    S.C = p.hidden << s;
    S.df = p.df[s];
    res /= S.df;
    S.Hs = res;
    S.heads = p.heads[s];
    S.nblocks = p.layers[s];
    S.Kp = c_in * S.df * S.df;
    S.M = 1LL * p.B * res * res;
    S.wp = add_param(p, 1LL * S.C * S.Kp);
    S.bp = add_param(p, S.C);
    S.wp16 = add_w16(p, 1LL * S.C * S.Kp);
    S.wp16t = add_w16(p, 1LL * S.C * S.Kp);
    const int C = S.C;
    for (int b = 0; b < S.nblocks; ++b) {
      BlockParams q;
      q.ln1_w = add_param(p, C); q.ln1_b = add_param(p, C);
      q.pos = add_param(p, (2 * kWindow - 1) * (2 * kWindow - 1));
      q.wqkv = add_param(p, 3LL * C * C);
      q.wo = add_param(p, 1LL * C * C); q.bo = add_param(p, C);
      q.ln2_w = add_param(p, C); q.ln2_b = add_param(p, C);
      q.w1 = add_param(p, 4LL * C * C); q.b1 = add_param(p, 4LL * C);
      q.w2 = add_param(p, 4LL * C * C); q.b2 = add_param(p, C);
      q.wqkv16 = add_w16(p, 3LL * C * C); q.wqkv16t = add_w16(p, 3LL * C * C);
      q.wo16 = add_w16(p, 1LL * C * C); q.wo16t = add_w16(p, 1LL * C * C);
      q.w116 = add_w16(p, 4LL * C * C); q.w116t = add_w16(p, 4LL * C * C);
      q.w216 = add_w16(p, 4LL * C * C); q.w216t = add_w16(p, 4LL * C * C);
      S.bp_.push_back(q);
    }
    c_in = C;
  }
  const int C4 = p.st[3].C;
  p.head_ln_w = add_param(p, C4); p.head_ln_b = add_param(p, C4);
  p.head_w = add_param(p, 1LL * p.num_classes * C4); p.head_b = add_param(p, p.num_classes);
  p.head_w16 = add_w16(p, 1LL * p.num_classes * C4); p.head_w16t = add_w16(p, 1LL * p.num_classes * C4);

  // ---- workspace.  Training keeps every block's activations; inference shares one block's worth.
  long long maxMC = 0;
  for (int s = 0; s < 4; ++s) {
    Stage& S = p.st[s];
    const long long M = S.M; const int C = S.C;
    maxMC = std::max(maxMC, M * C);
    S.cols = add_ws(p, M * S.Kp * 2);
    S.x0 = add_ws(p, M * C * 2);
    for (int b = 0; b < S.nblocks; ++b) {
      BlockActs a;
      if (p.training || (s == 0 && b == 0)) {
        // inference: one block's worth of buffers, sized by stage 1 (largest M*C), reused by every block
        const long long Mx = p.training ? M : p.st[0].M;
        const long long mc = p.training ? M * C : p.st[0].M * p.st[0].C;
        a.mean1 = add_ws(p, Mx * 4); a.rstd1 = add_ws(p, Mx * 4);
        a.xn1 = add_ws(p, mc * 2);
        a.qkv = add_ws(p, mc * 3 * 2);
        a.attn = add_ws(p, mc * 2);
        a.lse = add_ws(p, p.training ? M * S.heads * 4 : 256);
        a.xmid = add_ws(p, mc * 2);
        a.mean2 = add_ws(p, Mx * 4); a.rstd2 = add_ws(p, Mx * 4);
        a.xn2 = add_ws(p, mc * 2);
        a.hgrad = add_ws(p, p.training ? mc * 4 * 2 : 256);
        a.hact = add_ws(p, mc * 4 * 2);
        a.xout = add_ws(p, mc * 2);
      } else {
        a = p.st[0].ba_[0];
      }
      S.ba_.push_back(a);
    }
  }
  p.pooled = add_ws(p, 1LL * p.B * C4 * 2);
  p.pool_mean = add_ws(p, 1LL * p.B * 4); p.pool_rstd = add_ws(p, 1LL * p.B * 4);
  p.pooled_n = add_ws(p, 1LL * p.B * C4 * 2);
  if (p.training) {
    // the widest operand of any GEMM in a stage has 4C columns (stage 1 patch: Kp=48 < 4C)
    long long big = 0;
    for (int s = 0; s < 4; ++s) {
      const Stage& S = p.st[s];
      const long long wide = std::max<long long>(4LL * S.C, S.Kp);
      big = std::max(big, align_up(S.M, 8) * wide);
    }
    p.g_a = add_ws(p, maxMC * 2); p.g_b = add_ws(p, maxMC * 2);
    p.d_big = add_ws(p, big * 2);
    p.d_small = add_ws(p, maxMC * 2);
    // Partial sums of every fixed-order reduction of one Swin block's backward (4 split-K weight gradients of at most
    // ~#SMs tiles of 128 x 256 fp32 each, the LayerNorm / bias-gradient partial rows, the attention scratch): each gets its
    // own slice, the block's reductions are folded by ONE batched launch, then the arena is reused.
    p.arena_bytes = 4 * (1LL * 160 * 128 * 256 * 4) + 2 * (1LL * 8 * 160 * 3 * 1536 * 4) + 1LL * 2 * 160 * 4 * 1536 * 4 +
                    b200_window_attn_bwd_scratch_floats(b200_num_sms()) * 4 + (1 << 20);
    p.arena = add_ws(p, 2 * p.arena_bytes);      // two halves: a block's reductions (side stream) run while the next block fills the other
    p.demb16 = add_ws(p, 1LL * p.B * p.num_classes * 2);
    p.dpooled = add_ws(p, 1LL * p.B * C4 * 2);
  }
}

// ---------------------------------------------------------------------------------------------
struct Ctx {
  const Plan& p;
  const float* params; float* grads; bf16* wc; uint8_t* ws; cudaStream_t st; void* stv;
  cudaStream_t side = nullptr;           // null: no second stream
  mutable bool side_busy = false;        // the side stream holds work the main stream has not waited for yet
  mutable long long arena_used = 0;
  // what the side stream launches next may read everything the main stream has been given so far
  int fork() const {
    if (side == nullptr) return B200_OK;
    B200_CHECK_CUDA(cudaEventRecord(p.ev_fork, st));
    B200_CHECK_CUDA(cudaStreamWaitEvent(side, p.ev_fork, 0));
    return B200_OK;
  }
  // what the main stream launches next may overwrite the inputs / read the outputs of everything given to the side stream
  int join() const {
    if (side == nullptr || !side_busy) return B200_OK;
    B200_CHECK_CUDA(cudaEventRecord(p.ev_join, side));
    B200_CHECK_CUDA(cudaStreamWaitEvent(st, p.ev_join, 0));
    side_busy = false;
    return B200_OK;
  }
  void* wgrad_stream() const { if (side != nullptr) side_busy = true; return side != nullptr ? static_cast<void*>(side) : stv; }
  // slot = which of a block's four weight gradients was just given to the side stream / must have finished before the main
  // stream overwrites one of its operands
  int done(int slot) const {
    if (side == nullptr) return B200_OK;
    B200_CHECK_CUDA(cudaEventRecord(p.ev_done[slot], side));
    return B200_OK;
  }
  int wait(int slot) const {
    if (side == nullptr) return B200_OK;
    B200_CHECK_CUDA(cudaStreamWaitEvent(st, p.ev_done[slot], 0));
    return B200_OK;
  }
  // a slice of the partial-sum arena; when it runs out the recorded reductions are folded and the arena starts over
  float* take(long long bytes, int* rc) const {
    bytes = align_up(bytes, 256);
    if (arena_used + bytes > p.arena_bytes) {
      *rc = bytes > p.arena_bytes ? B200_ERR_INVALID : flush(true);
      if (*rc) return nullptr;
    }
    float* ptr = reinterpret_cast<float*>(ws + p.arena + (side != nullptr ? p.half * p.arena_bytes : 0) + arena_used);
    arena_used += bytes;
    return ptr;
  }
  int flush(bool keep) const {
    arena_used = 0;
    if (side == nullptr) return b200_reduce_flush(stv, keep ? 1 : 0);
    // Two streams: the batched reduction follows the weight gradients on the side stream (it also reads partial sums the main
    // stream produced: LayerNorm rows, the attention scratch - hence the fork) and the main stream moves on to the next block,
    // which fills the other half of the arena; a half is written again only after the reduction that read it has run.
    int rc = fork();
    if (rc) return rc;
    rc = b200_reduce_flush(side, keep ? 1 : 0);
    if (rc) return rc;
    B200_CHECK_CUDA(cudaEventRecord(p.ev_red[p.half], side));
    side_busy = true;
    p.half ^= 1;
    B200_CHECK_CUDA(cudaStreamWaitEvent(st, p.ev_red[p.half], 0));
    return B200_OK;
  }
  template <class T> T* W(long long off) const { return reinterpret_cast<T*>(ws + off); }
  const float* P(const ParamRef& r) const { return params + r.off; }
  float* G(const ParamRef& r) const { return grads + r.off; }
};

#define RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

int linear_fwd(const Ctx& c, const bf16* x, long long M, int K, const bf16* w16, int N, const float* bias, int mode, bf16* out,
               bf16* out2, const bf16* aux) {
  return b200_gemm_tn(x, K, w16, K, static_cast<int>(M), N, K, 1, mode, out, N, 0, out2, N, bias, aux, N, 1, 0, 0, c.stv);
}

// dX[M,K] = dY[M,N] * W[N,K]  ==  TN GEMM with B = W^T [K,N]
int linear_dgrad(const Ctx& c, const bf16* dy, long long M, int N, const bf16* w16t, int K, int mode, bf16* dx, const bf16* aux) {
  return b200_gemm_tn(dy, N, w16t, N, static_cast<int>(M), K, N, 1, mode, dx, K, 0, nullptr, 0, nullptr, aux, K, 1, 0, 0, c.stv);
}

// dW[N,K] = dY[M,N]^T * X[M,K]: operands read in place (MN-major tcgen05 descriptors), split-K over the tokens,
// fp32 partials reduced in a fixed order.
int bias_grad(const Ctx& c, const bf16* dy, long long M, int N, float* db, void* stv);
// db (optional): the bias gradient colsum(dy), produced by the same kernel when the tile shape allows it.
// The GEMM goes to the side stream when there is one (call c.fork() where its inputs are final and c.join() before they are
// overwritten); the reductions of its partial sums are recorded as always and run at the block's flush, after the join.
int linear_wgrad(const Ctx& c, const bf16* dy, long long M, int N, const bf16* x, int K, float* dw, float* db = nullptr) {
  const int bn = (K <= 256) ? static_cast<int>(align_up(K, 16)) : 0;
  const int bn_eff = bn ? bn : 256;
  const long long tiles = ((N + 127) / 128) * ((K + bn_eff - 1) / bn_eff);
  int splits = static_cast<int>(std::max<long long>(1, b200_num_sms() / tiles));
  splits = b200_gemm_splits(static_cast<int>(M), splits);
  while (splits > 1 && 1LL * splits * N * K * 4 > c.p.arena_bytes / 4) --splits;
  splits = b200_gemm_splits(static_cast<int>(M), splits);
  int rc = 0;
  float* partial = c.take(1LL * splits * N * K * 4, &rc);
  RC(rc);
  float* cs = nullptr;
  if (db != nullptr) {
    cs = c.take(1LL * splits * N * 4, &rc);
    RC(rc);
  }
  void* wst = c.wgrad_stream();          // after the takes: a take may flush, and a flush joins
  if (db == nullptr) {
    RC(b200_gemm_wgrad(dy, N, x, K, M, N, K, partial, splits, bn, wst));
    return b200_splitk_reduce(partial, dw, 1LL * N * K, splits, 0, c.stv);
  }
  int fused = 0;
  RC(b200_gemm_wgrad_bias(dy, N, x, K, M, N, K, partial, cs, splits, bn, &fused, wst));
  RC(b200_splitk_reduce(partial, dw, 1LL * N * K, splits, 0, c.stv));
  if (fused) return b200_splitk_reduce(cs, db, N, splits, 0, c.stv);
  return bias_grad(c, dy, M, N, db, wst);
}

int bias_grad(const Ctx& c, const bf16* dy, long long M, int N, float* db, void* stv) {
  int rc = 0;
  float* partial = c.take(1LL * b200_colsum_blocks(M) * N * 4, &rc);
  RC(rc);
  return b200_colsum(dy, N, M, N, db, partial, 0, stv);
}

// LayerNorm backward with its partial rows in the arena.  win_shift >= 0: dy is in window-major order (LN1 of a block whose
// shift flag is win_shift)
int ln_bwd_win(const Ctx& c, const bf16* dy, const bf16* x, const float* gamma, const float* mean, const float* rstd, const bf16* dres,
               bf16* dx, float* dgamma, float* dbeta, float* dres_colsum, int B, int Hs, int C, int shifted) {
  int rc = 0;
  float* partial = c.take(1LL * b200_layernorm_bwd_blocks(1LL * B * Hs * Hs, C) * 3 * C * 4, &rc);
  RC(rc);
  return b200_layernorm_bwd_windows(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, dres_colsum, partial, B, Hs, Hs, C, shifted, 0, c.stv);
}
int ln_bwd(const Ctx& c, const bf16* dy, const bf16* x, const float* gamma, const float* mean, const float* rstd, const bf16* dres,
           bf16* dx, float* dgamma, float* dbeta, float* dres_colsum, long long M, int C) {
  int rc = 0;
  float* partial = c.take(1LL * b200_layernorm_bwd_blocks(M, C) * 3 * C * 4, &rc);
  RC(rc);
  return b200_layernorm_bwd(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, dres_colsum, partial, M, C, 0, c.stv);
}

int forward(const Ctx& c, const void* img, int img_u8, float* emb) {
  const Plan& p = c.p;
  const bf16* x = nullptr;
  for (int s = 0; s < 4; ++s) {
    const Stage& S = p.st[s];
    bf16* cols = c.W<bf16>(S.cols);
    if (s == 0) {
      if (img_u8) RC(b200_patch_gather_image_u8(reinterpret_cast<const unsigned char*>(img), cols, p.B, p.channels, p.img, p.img, S.df, S.Kp, c.stv));
      else RC(b200_patch_gather_image(reinterpret_cast<const float*>(img), cols, p.B, p.channels, p.img, p.img, S.df, S.Kp, c.stv));
    }
    else RC(b200_patch_gather_nhwc(const_cast<bf16*>(x), cols, p.B, p.st[s - 1].Hs, p.st[s - 1].Hs, p.st[s - 1].C, 0, c.stv));
    bf16* x0 = c.W<bf16>(S.x0);
    RC(linear_fwd(c, cols, S.M, S.Kp, c.wc + S.wp16, S.C, c.P(S.bp), B200_EPI_STORE, x0, nullptr, nullptr));
    x = x0;
    const int C = S.C;
    for (int b = 0; b < S.nblocks; ++b) {
      const BlockParams& q = S.bp_[b];
      const BlockActs& a = S.ba_[b];
      bf16* xn1 = c.W<bf16>(a.xn1);
      // LN1 writes its rows in window-major order of this block's (shifted) partition: xn1, qkv, lse, dqkv and d xn1 all live
      // in that order (the GEMMs between them are row-wise); the attention output comes back in raster order
      RC(b200_layernorm_fwd_windows(x, c.P(q.ln1_w), c.P(q.ln1_b), xn1, c.W<float>(a.mean1), c.W<float>(a.rstd1), p.B, S.Hs, S.Hs, C, b & 1,
                                    1e-5f, c.stv));
      bf16* qkv = c.W<bf16>(a.qkv);
      RC(linear_fwd(c, xn1, S.M, C, c.wc + q.wqkv16, 3 * C, nullptr, B200_EPI_STORE, qkv, nullptr, nullptr));
      bf16* attn = c.W<bf16>(a.attn);
      RC(b200_window_attn_fwd(qkv, c.P(q.pos), attn, p.training ? c.W<float>(a.lse) : nullptr, p.B, S.Hs, S.Hs, C, S.heads, b & 1, c.stv));
      bf16* xmid = c.W<bf16>(a.xmid);
      RC(linear_fwd(c, attn, S.M, C, c.wc + q.wo16, C, c.P(q.bo), B200_EPI_RESID, xmid, nullptr, x));
      bf16* xn2 = c.W<bf16>(a.xn2);
      RC(b200_layernorm_fwd(xmid, c.P(q.ln2_w), c.P(q.ln2_b), xn2, c.W<float>(a.mean2), c.W<float>(a.rstd2), S.M, C, 1e-5f, c.stv));
      bf16* hact = c.W<bf16>(a.hact);
      // training also keeps GELU'(pre-activation) (out2): backward multiplies by it
      RC(linear_fwd(c, xn2, S.M, C, c.wc + q.w116, 4 * C, c.P(q.b1), B200_EPI_GELU, hact, p.training ? c.W<bf16>(a.hgrad) : nullptr, nullptr));
      // inference shares one block's buffers: the block input may live in xout, so alternate with x0
      bf16* xout = c.W<bf16>(a.xout);
      if (!p.training && x == xout) xout = c.W<bf16>(p.st[0].x0);
      RC(linear_fwd(c, hact, S.M, 4 * C, c.wc + q.w216, C, c.P(q.b2), B200_EPI_RESID, xout, nullptr, xmid));
      x = xout;
    }
  }
  const Stage& L = p.st[3];
  bf16* pooled = c.W<bf16>(p.pooled);
  RC(b200_mean_pool(x, pooled, p.B, L.Hs * L.Hs, L.C, 0, c.stv));
  bf16* pn = c.W<bf16>(p.pooled_n);
  RC(b200_layernorm_fwd(pooled, c.P(p.head_ln_w), c.P(p.head_ln_b), pn, c.W<float>(p.pool_mean), c.W<float>(p.pool_rstd), p.B, L.C, 1e-5f, c.stv));
  return b200_gemm_tn(pn, L.C, c.wc + p.head_w16, L.C, p.B, p.num_classes, L.C, 1, B200_EPI_STORE, emb, p.num_classes, 1, nullptr, 0,
                      c.P(p.head_b), nullptr, 0, 1, 0, 0, c.stv);
}

// stage_lo..stage_hi (inclusive, descending) lets the host interleave gradient all-reduce buckets
int backward(const Ctx& c, const float* demb, int stage_hi, int stage_lo) {
  const Plan& p = c.p;
  const Stage& L = p.st[3];
  // gradient of the residual stream entering stage s lives in g_a for s = 3, 1 and in g_b for s = 2, 0
  // (fixed by parity so that a backward split into several stage ranges finds it again)
  auto gbuf = [&](int s) { return c.W<bf16>(((3 - s) & 1) ? p.g_b : p.g_a); };
  if (stage_hi >= 4) {
    bf16* g = gbuf(3);
    // ---- head: Linear <- LayerNorm <- mean pool  (models/swin.py:214-217, :224)
    bf16* d16 = c.W<bf16>(p.demb16);
    RC(b200_cast_f32_bf16(demb, d16, 1LL * p.B * p.num_classes, c.stv));
    RC(c.fork());                                            // the head's weight gradient (side stream) reads d16
    RC(linear_wgrad(c, d16, p.B, p.num_classes, c.W<bf16>(p.pooled_n), L.C, c.G(p.head_w)));
    RC(bias_grad(c, d16, p.B, p.num_classes, c.G(p.head_b), c.stv));
    bf16* dpn = c.W<bf16>(p.dpooled);
    RC(linear_dgrad(c, d16, p.B, p.num_classes, c.wc + p.head_w16t, L.C, B200_EPI_STORE, dpn, nullptr));
    bf16* dpool = c.W<bf16>(p.d_small);
    RC(ln_bwd(c, dpn, c.W<bf16>(p.pooled), c.P(p.head_ln_w), c.W<float>(p.pool_mean), c.W<float>(p.pool_rstd), nullptr, dpool,
              c.G(p.head_ln_w), c.G(p.head_ln_b), nullptr, p.B, L.C));
    RC(b200_mean_pool(dpool, g, p.B, L.Hs * L.Hs, L.C, 1, c.stv));
    stage_hi = 3;
  }
  for (int s = stage_hi; s >= stage_lo; --s) {
    const Stage& S = p.st[s];
    const int C = S.C;
    const long long M = S.M;
    bf16* g = gbuf(s);
    for (int b = S.nblocks - 1; b >= 0; --b) {
      const BlockParams& q = S.bp_[b];
      const BlockActs& a = S.ba_[b];
      const bf16* x_in = (b == 0) ? c.W<bf16>(S.x0) : c.W<bf16>(S.ba_[b - 1].xout);
      bf16* dbig = c.W<bf16>(p.d_big);
      bf16* dsmall = c.W<bf16>(p.d_small);
      // ---- MLP: x_out = x_mid + W2 gelu(W1 LN2(x_mid) + b1) + b2
      // The four weight gradients of a block go to the side stream (see Plan::side): fork() marks the point from which their
      // operands are final, done(i) / wait(i) keep the main stream from overwriting an operand (g, dbig) before they have read it.
      RC(c.fork());
      RC(c.wait(3));                                         // the previous block's qkv weight gradient has read d_big
      RC(linear_dgrad(c, g, M, C, c.wc + q.w216t, 4 * C, B200_EPI_DGELU, dbig, c.W<bf16>(a.hgrad)));   // d h_pre = (dy W2) o gelu'(h_pre)
      RC(linear_wgrad(c, g, M, C, c.W<bf16>(a.hact), 4 * C, c.G(q.w2)));
      RC(c.done(0));
      RC(c.fork());
      RC(linear_dgrad(c, dbig, M, 4 * C, c.wc + q.w116t, C, B200_EPI_STORE, dsmall, nullptr));        // d xn2
      RC(linear_wgrad(c, dbig, M, 4 * C, c.W<bf16>(a.xn2), C, c.G(q.w1), c.G(q.b1)));            // + d b1 = colsum(d h_pre)
      RC(c.done(1));
      // g <- d x_mid; the same pass yields d b2 = colsum(g): g is dL/d(fc2 output)
      RC(c.wait(0));
      RC(ln_bwd(c, dsmall, c.W<bf16>(a.xmid), c.P(q.ln2_w), c.W<float>(a.mean2), c.W<float>(a.rstd2), g, g,
                c.G(q.ln2_w), c.G(q.ln2_b), c.G(q.b2), M, C));
      // ---- attention: x_mid = x_in + Wo attn(LN1(x_in)) + bo
      RC(c.fork());
      RC(linear_dgrad(c, g, M, C, c.wc + q.wo16t, C, B200_EPI_STORE, dsmall, nullptr));               // d attn_out
      RC(linear_wgrad(c, g, M, C, c.W<bf16>(a.attn), C, c.G(q.wo)));
      RC(c.done(2));
      {
        int rc = 0;
        float* scratch = c.take(b200_window_attn_bwd_scratch_floats(b200_window_attn_bwd_blocks(p.B, S.Hs, S.Hs, S.heads)) * 4, &rc);
        RC(rc);
        RC(c.wait(1));
        RC(b200_window_attn_bwd(c.W<bf16>(a.qkv), c.P(q.pos), c.W<float>(a.lse), dsmall, dbig, c.G(q.pos), scratch, 0, p.B, S.Hs, S.Hs, C,
                                S.heads, b & 1, c.stv));                                                // dbig <- d qkv
      }
      RC(c.fork());
      RC(linear_dgrad(c, dbig, M, 3 * C, c.wc + q.wqkv16t, C, B200_EPI_STORE, dsmall, nullptr));      // d xn1
      RC(linear_wgrad(c, dbig, M, 3 * C, c.W<bf16>(a.xn1), C, c.G(q.wqkv)));
      RC(c.done(3));
      // g <- d x_in; d bo = colsum(g before the update): g is dL/d(to_out output)
      RC(c.wait(2));
      RC(ln_bwd_win(c, dsmall, x_in, c.P(q.ln1_w), c.W<float>(a.mean1), c.W<float>(a.rstd1), g, g, c.G(q.ln1_w), c.G(q.ln1_b), c.G(q.bo),
                    p.B, S.Hs, C, b & 1));
      RC(c.flush(true));                                     // the block's eight reductions: one launch
    }
    // ---- patch merging linear (models/swin.py:162-167)
    RC(c.fork());
    RC(linear_wgrad(c, g, M, C, c.W<bf16>(S.cols), S.Kp, c.G(S.wp), c.G(S.bp)));
    if (s > 0) {
      bf16* dcols = c.W<bf16>(p.d_big);
      RC(c.wait(3));                                         // the last block's qkv weight gradient has read d_big
      RC(linear_dgrad(c, g, M, C, c.wc + S.wp16t, S.Kp, B200_EPI_STORE, dcols, nullptr));
      RC(b200_patch_gather_nhwc(gbuf(s - 1), dcols, p.B, p.st[s - 1].Hs, p.st[s - 1].Hs, p.st[s - 1].C, 1, c.stv));
    }
    RC(c.join());                                            // the patch-merging weight gradient reads g, which the next stage reuses
  }
  return c.join();
}

// fp32 master weights -> the bf16 (and transposed bf16) copies the GEMMs read: one launch over a table of all matrices
int sync_weights(const Ctx& c) {
  const Plan& p = c.p;
  CastJobs jobs{};
  int tiles = 0;
  int rc = B200_OK;
  // one launch per table of 64 matrices: Swin-T (49) takes one, the 24-block variants (Swin-S / B / L: 101) two
  auto add = [&](long long in_off, long long dst_off, long long dst_t_off, int R, int Cc) {
    if (rc) return;
    if (jobs.n == 64) {
      rc = cast_transpose_multi(c.params, c.wc, jobs, tiles, c.st);
      jobs.n = 0;
      tiles = 0;
    }
    CastJob& j = jobs.job[jobs.n++];
    j.in_off = in_off; j.dst_off = dst_off; j.dst_t_off = dst_t_off; j.R = R; j.Cc = Cc;
    j.tile0 = tiles; j.tiles_c = (Cc + 31) / 32;
    tiles += ((R + 31) / 32) * j.tiles_c;
  };
  for (int s = 0; s < 4; ++s) {
    const Stage& S = p.st[s];
    add(S.wp.off, S.wp16, s > 0 ? S.wp16t : -1, S.C, S.Kp);
    const int C = S.C;
    for (const BlockParams& q : S.bp_) {
      add(q.wqkv.off, q.wqkv16, q.wqkv16t, 3 * C, C);
      add(q.wo.off, q.wo16, q.wo16t, C, C);
      add(q.w1.off, q.w116, q.w116t, 4 * C, C);
      add(q.w2.off, q.w216, q.w216t, C, 4 * C);
    }
  }
  add(p.head_w.off, p.head_w16, p.head_w16t, p.num_classes, p.st[3].C);
  RC(rc);
  return cast_transpose_multi(c.params, c.wc, jobs, tiles, c.st);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" void* b200_swin_create(int batch, int img, int channels, int hidden_dim, const int* layers, const int* heads,
                                  const int* downscaling, int num_classes, int head_dim, int window_size, int training) {
  if (batch <= 0 || img <= 0 || channels <= 0 || hidden_dim <= 0 || num_classes <= 0) {
    b200_set_error(B200_ERR_INVALID, "swin_create: non-positive dimension");
    return nullptr;
  }
  if (head_dim != kHeadDim || window_size != kWindow) {
    b200_set_error(B200_ERR_INVALID, "swin_create: only head_dim=32 / window_size=7 (the reference's fixed values) are built");
    return nullptr;
  }
  int res = img, c_in = channels;
  for (int s = 0; s < 4; ++s) {
    const int df = downscaling[s];
    if ((s == 0 && df != 4) || (s > 0 && df != 2) || res % df != 0 || (res / df) % kWindow != 0 || layers[s] <= 0 || layers[s] % 2 != 0 ||
        heads[s] * kHeadDim != (hidden_dim << s) || (hidden_dim << s) % 32 != 0 || (hidden_dim << s) > 1536) {
      b200_set_error(B200_ERR_INVALID, "swin_create: stage %d unsupported (df=%d res=%d layers=%d heads=%d C=%d)", s, df, res, layers[s],
                     heads[s], hidden_dim << s);
      return nullptr;
    }
    res /= df;
    c_in = hidden_dim << s;
  }
  (void)c_in;
  if (num_classes % 8 != 0) {
    b200_set_error(B200_ERR_INVALID, "swin_create: num_classes must be a multiple of 8");
    return nullptr;
  }
  Plan* p = new Plan();
  p->B = batch; p->img = img; p->channels = channels; p->hidden = hidden_dim; p->num_classes = num_classes; p->training = training;
  for (int s = 0; s < 4; ++s) { p->layers[s] = layers[s]; p->heads[s] = heads[s]; p->df[s] = downscaling[s]; }
  build(*p);
  return p;
}

extern "C" void b200_swin_destroy(void* plan) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (p == nullptr) return;
  if (p->side != nullptr) {
    cudaStreamSynchronize(p->side);
    cudaStreamDestroy(p->side);
    cudaEventDestroy(p->ev_fork);
    cudaEventDestroy(p->ev_join);
    for (auto& e : p->ev_done) cudaEventDestroy(e);
    for (auto& e : p->ev_red) cudaEventDestroy(e);
  }
  delete p;
}

extern "C" long long b200_swin_param_elems(const void* plan) { return reinterpret_cast<const Plan*>(plan)->n_params; }
extern "C" int b200_swin_param_count(const void* plan) { return static_cast<int>(reinterpret_cast<const Plan*>(plan)->order.size()); }
extern "C" int b200_swin_param_offsets(const void* plan, long long* offsets, long long* numels, int n) {
  const Plan* p = reinterpret_cast<const Plan*>(plan);
  B200_REQUIRE(n == static_cast<int>(p->order.size()), "swin_param_offsets: expected %d entries", (int)p->order.size());
  for (int i = 0; i < n; ++i) { offsets[i] = p->order[i].off; numels[i] = p->order[i].numel; }
  return B200_OK;
}
extern "C" long long b200_swin_wcache_bytes(const void* plan) { return reinterpret_cast<const Plan*>(plan)->wcache_elems * 2; }
extern "C" long long b200_swin_workspace_bytes(const void* plan) { return reinterpret_cast<const Plan*>(plan)->ws_bytes; }

extern "C" int b200_swin_sync_weights(const void* plan, const float* params, void* wcache, void* stream) {
  const Plan* p = reinterpret_cast<const Plan*>(plan);
  Ctx c{*p, params, nullptr, reinterpret_cast<bf16*>(wcache), nullptr, reinterpret_cast<cudaStream_t>(stream), stream};
  return sync_weights(c);
}

extern "C" int b200_swin_forward(const void* plan, const float* params, const void* wcache, const void* img, int img_is_u8, float* emb,
                                 void* workspace, long long workspace_bytes, void* stream) {
  const Plan* p = reinterpret_cast<const Plan*>(plan);
  if (workspace_bytes < p->ws_bytes)
    return b200_set_error(B200_ERR_WORKSPACE, "swin_forward: workspace %lld < required %lld bytes", workspace_bytes, p->ws_bytes);
  Ctx c{*p, params, nullptr, reinterpret_cast<bf16*>(const_cast<void*>(wcache)), reinterpret_cast<uint8_t*>(workspace),
        reinterpret_cast<cudaStream_t>(stream), stream};
  return forward(c, img, img_is_u8, emb);
}

extern "C" int b200_swin_backward(const void* plan, const float* params, const void* wcache, const float* demb, float* grads,
                                  void* workspace, long long workspace_bytes, int stage_hi, int stage_lo, void* stream) {
  const Plan* p = reinterpret_cast<const Plan*>(plan);
  B200_REQUIRE(p->training, "swin_backward: plan was created for inference");
  B200_REQUIRE(stage_hi >= stage_lo && stage_lo >= 0 && stage_hi <= 4, "swin_backward: bad stage range [%d, %d]", stage_lo, stage_hi);
  if (workspace_bytes < p->ws_bytes)
    return b200_set_error(B200_ERR_WORKSPACE, "swin_backward: workspace %lld < required %lld bytes", workspace_bytes, p->ws_bytes);
  Ctx c{*p, params, grads, reinterpret_cast<bf16*>(const_cast<void*>(wcache)), reinterpret_cast<uint8_t*>(workspace),
        reinterpret_cast<cudaStream_t>(stream), stream};
  static const bool side_on = [] { const char* e = getenv("B200_WGRAD_STREAM"); return e == nullptr || e[0] != '0'; }();
  // while launches are being timed one by one (bench.py's roofline region) everything stays on one stream: a kernel's
  // event pair must bracket that kernel alone, not its competition with another stream for SMs
  if (side_on && !b200_prof_timing()) {
    if (p->side == nullptr) {
      B200_CHECK_CUDA(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
      B200_CHECK_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
      B200_CHECK_CUDA(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
      for (auto& e : p->ev_done) B200_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      for (auto& e : p->ev_red) B200_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    c.side = p->side;
  }
  // every fixed-order reduction of the pass is recorded and folded one batch per block (see Ctx::take / flush); whatever
  // happens, deferred mode ends with this call
  int rc = b200_reduce_defer_begin();
  if (rc == B200_OK) rc = backward(c, demb, stage_hi, stage_lo);
  const int rc_join = c.join();          // also on the error path: nothing may stay behind on the side stream
  const int rc_flush = b200_reduce_flush(stream, 0);
  return rc ? rc : (rc_join ? rc_join : rc_flush);
}
