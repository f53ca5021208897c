/* b200_fe.h - C ABI of libb200fe.so: the B200 (sm_100a) replacement for the FE-training and
 * gallery-matching hot path of MarQuisCheshire/Pets-Face-Recognition.
 *
 * The reference has no FFI of its own: the path sits behind a Python plugin API (config.model(),
 * config.loss(), config.similarity_f(), engine.Controller / engine.Trainer - SURVEY.md section 8b) and every
 * FLOP is a stock PyTorch operator.  This header is therefore the boundary a maintainer binds with
 * ctypes from those Python hooks (INTEGRATION.md shows the stubs); each entry cites the reference
 * lines whose arithmetic it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or a negative B200_ERR_* code; b200_last_error() gives the
 *     thread-local message.  Nothing throws, allocates or frees caller memory, or synchronises the device.
 *   - pointers are raw device pointers unless noted; `stream` is a cudaStream_t (void*).
 *   - bf16 / fp16 tensors are passed as void*; "ld*" are row pitches in ELEMENTS; int64 labels.
 *   - workspace sizes come from the *_bytes / *_blocks query functions; the caller allocates.
 */
#ifndef B200_FE_H_
#define B200_FE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ERR_INVALID (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_WORKSPACE (-3)
#define B200_ERR_ARCH (-4)

/* epilogues of b200_gemm_tn */
#define B200_EPI_STORE 0   /* out = acc (+ bias)                                   nn.Linear                  */
#define B200_EPI_GELU 1    /* h = acc + bias, out = gelu_erf(h), out2 (optional) = gelu_erf'(h)   models/swin.py:39-41 */
#define B200_EPI_RESID 2   /* out = acc + bias + aux                               Residual, models/swin.py:22-23 */
#define B200_EPI_DGELU 3   /* out = acc * aux, aux = the out2 saved by B200_EPI_GELU          autograd of models/swin.py:41 */
#define B200_EPI_PARTIAL 4 /* fp32 out[split] = acc   (split-K partial, weight gradients)                     */
#define B200_EPI_GELU_Q8 5  /* B200_EPI_GELU with out2 stored as 8-bit codes: uint8 [M, N], q = rint((g + 0.13) * 255 / 1.26) */
#define B200_EPI_DGELU_Q8 6 /* B200_EPI_DGELU with aux = those uint8 codes                                     */

/* b200_cosine_topk: pass as exclude_self_offset when no gallery row is to be skipped */
#define B200_NO_EXCLUDE (-(1LL << 62))

#define B200_OPT_SGD 0
#define B200_OPT_ADAMW 1

/* one parameter tensor of a fused optimizer step (device-resident table, see b200_optimizer_step) */
typedef struct B200OptTensor {
  void* param;       /* fp32 [numel], updated in place                                  */
  const void* grad;  /* fp32 [numel]                                                     */
  void* state1;      /* SGD: momentum buffer; AdamW: exp_avg                             */
  void* state2;      /* AdamW: exp_avg_sq (unused for SGD)                               */
  void* param_bf16;  /* optional bf16 copy refreshed by the same kernel, or NULL         */
  long long numel;
  float lr, weight_decay, beta1 /* SGD: momentum */, beta2, eps;
  int step;          /* steps already taken (0 = first step: SGD buffer := grad)         */
} B200OptTensor;

const char* b200_last_error(void);
int b200_device_check(void); /* B200_ERR_ARCH unless the current device is sm_100 */

/* ---- measurement hooks (bench.py): count the library's kernel launches; optionally bracket every tcgen05 GEMM launch
 *      with CUDA events on its stream.  b200_prof_end synchronises the device. ------------------------------------ */
int b200_prof_begin(int time_gemm_launches);
int b200_prof_end(double* gemm_ms, double* gemm_flops, long long* gemm_launches, long long* total_launches);
int b200_prof_gemm_bytes(double* bytes); /* algorithmic HBM bytes of the launches the last b200_prof_end summed */
/* kinds of timed kernels: the tcgen05 GEMM, window attention forward / backward, LayerNorm forward / backward, the batched
 * fixed-order reductions */
#define B200_PROF_GEMM 0
#define B200_PROF_ATTN_FWD 1
#define B200_PROF_ATTN_BWD 2
#define B200_PROF_LN_FWD 3
#define B200_PROF_LN_BWD 4
#define B200_PROF_REDUCE 5
#define B200_PROF_KINDS 6
/* per-kind totals (CUDA-event ms, algorithmic FLOPs and HBM bytes, launches) of the last timed region; call after b200_prof_end */
int b200_prof_kernels(int n_kinds, double* ms, double* flops, double* bytes, long long* launches);

/* ---- linear layers: D[M,N] = A[M,K] * B[N,K]^T on tcgen05 tensor cores (TMA-fed, TMEM accumulators) ---------
 * replaces nn.Linear forward and the two GEMMs autograd runs for its backward (models/swin.py:39-43,91,98,
 * 160,166,216; "Linear" row of SURVEY.md appendix B).  A and B are 16-bit (is_bf16 ? bf16 : fp16), K-major. */
int b200_gemm_tn(const void* a, long long lda, const void* b, long long ldb, int M, int N, int K, int is_bf16, int mode,
                 void* out, long long ldo, int out_fp32, void* out2, long long ldo2, const float* bias, const void* aux,
                 long long ldaux, int splits, long long split_stride, int block_n, void* stream);
/* weight gradient dW[N,K] = dY[tokens,N]^T * X[tokens,K] (bf16), split over the tokens into fp32 partials [splits][N][K];
 * operands are consumed in place through MN-major UMMA descriptors (no transposed copies) */
int b200_gemm_wgrad(const void* dy, long long ldy, const void* x, long long ldx, long long tokens, int N, int K, float* partial,
                    int splits, int block_n, void* stream);
/* b200_gemm_wgrad + the bias gradient colsum(dy) as [splits][N] fp32 partial rows (all-ones MMA inside the same kernel);
 * *fused = 0 when the tile shape has no room for it (then colsum_partial is untouched: use b200_colsum). */
int b200_gemm_wgrad_bias(const void* dy, long long ldy, const void* x, long long ldx, long long tokens, int N, int K, float* partial,
                         float* colsum_partial, int splits, int block_n, int* fused, void* stream);
int b200_gemm_splits(int K, int splits); /* split count b200_gemm_tn will really use (sizes the partial buffer) */
int b200_splitk_reduce(const float* partial, float* out, long long n, int splits, int accumulate, void* stream);
/* Batched reductions.  Between b200_reduce_defer_begin() and b200_reduce_flush() (same host thread) every fixed-order
 * reduction the library would launch - b200_splitk_reduce and the ones inside b200_layernorm_bwd, b200_colsum and
 * b200_window_attn_bwd - is recorded instead; the flush folds all of them in ONE launch on `stream` (keep_deferring != 0
 * re-opens a batch).  Every recorded reduction needs its own partial buffer, alive until the flush; the outputs are
 * defined only after it.  The summation order of an output never depends on the rest of the batch. */
int b200_reduce_defer_begin(void);
int b200_reduce_pending(void);
int b200_reduce_flush(void* stream, int keep_deferring);

/* ---- LayerNorm, nn.LayerNorm(C) eps 1e-5 (models/swin.py:29,215) -------------------------------------------- */
int b200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, long long M,
                       int C, float eps, void* stream);
int b200_layernorm_bwd_blocks(long long M, int C); /* rows of the [blocks, 3C] fp32 partial buffer */
int b200_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                       const void* dres_in, void* dx_out, float* dgamma, float* dbeta, float* dres_colsum, float* partial,
                       long long M, int C, int accumulate, void* stream);

/* Window-major variants for the LayerNorm in front of to_qkv (models/swin.py:160 PreNorm -> WindowAttention :101-135): the
 * token grid [B, H, W] is cut into 7 x 7 windows (after the cyclic shift by 3 when shifted != 0, models/swin.py:8-14) and
 * window-major row = ((b * H/7 + wy) * W/7 + wx) * 49 + 7 r + c.  fwd writes y in that order; bwd reads dy in that order. */
int b200_layernorm_fwd_windows(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int B, int H,
                               int W, int C, int shifted, float eps, void* stream);
int b200_layernorm_bwd_windows(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                               const void* dres_in, void* dx_out, float* dgamma, float* dbeta, float* dres_colsum, float* partial,
                               int B, int H, int W, int C, int shifted, int accumulate, void* stream);
/* rows of `row_bytes` bytes (a multiple of 4) permuted between raster and window-major order: to_window != 0 writes
 * out[window_major(r)] = in[r], else out[r] = in[window_major(r)] */
int b200_window_rows(const void* in, void* out, int B, int H, int W, int row_bytes, int shifted, int to_window, void* stream);

/* ---- patch merging gather == nn.Unfold(k=s=df) + NHWC view (models/swin.py:159,162-165) ---------------------- */
int b200_patch_gather_image(const float* img_nchw, void* cols, int B, int Cin, int H, int W, int df, long long ldo, void* stream);
int b200_patch_gather_image_u8(const unsigned char* img_nchw, void* cols, int B, int Cin, int H, int W, int df, long long ldo,
                              void* stream); /* uint8 pixels, ToTensor's /255 fused */
int b200_patch_gather_nhwc(void* x_nhwc, void* cols, int B, int H, int W, int C, int backward, void* stream);

/* ---- training augmentation of configs/dog_fe/fe_dogs_config.py:17-26 on a uint8 batch in HBM: sharpness 0 (ImageFilter.SMOOTH),
 *      autocontrast, crop, PIL's two-pass fixed-point bilinear resize, nearest-neighbour rotation (fill 0) - PIL's own integer /
 *      float32 arithmetic, so equal random draws give equal bytes.  params_dev: one B200AugParams per image (device memory);
 *      coef_dev: int [S][4] = first source index + three 22-bit fixed-point bilinear weights of the crop -> S resize (16-B aligned);
 *      minmax_scratch: B * 6 bytes.  The result stays uint8: b200_swin_forward takes it as it is (x / 255 fused). ----------------- */
typedef struct B200AugParams {
  int sharpen;       /* RandomAdjustSharpness(0, p) fired  */
  int autocontrast;  /* RandomAutocontrast(p) fired        */
  int crop_y, crop_x;
  int rot[6];        /* PIL's affine walk of the rotation in 16.16 fixed point: x step, y step, origin for xin (0..2) and yin (3..5) */
} B200AugParams;
int b200_augment_train(const unsigned char* in_nchw, unsigned char* out_nchw, const B200AugParams* params_dev, const int* coef_dev, int B,
                       int H, int W, int crop, int S, unsigned char* minmax_scratch, void* stream);

/* ---- x.mean(dim=[2,3]) (models/swin.py:224) and its backward -------------------------------------------------- */
int b200_mean_pool(const void* in, void* out, int B, int T, int C, int backward, void* stream);

/* ---- layout / dtype helpers ------------------------------------------------------------------------------------ */
int b200_transpose16(const void* in, void* out, long long R, int Cc, long long ld_in, long long ld_out, void* stream);
int b200_cast_transpose(const float* in, void* dst_bf16, void* dst_t_bf16, int R, int Cc, void* stream);
int b200_cast_f32_bf16(const float* in, void* out, long long n, void* stream);
int b200_colsum_blocks(long long M);
int b200_colsum(const void* x, long long ld, long long M, int N, float* out, float* partial, int accumulate, void* stream);

/* ---- (shifted-)window attention (models/swin.py:101-135, CyclicShift :8-14, create_mask :49-62,
 *      get_relative_distances :65-68) on tcgen05 tensor cores, q / k / v tiles fetched by TMA.
 *      qkv: [B*H*W, 3C] bf16 = output of to_qkv with its rows in WINDOW-MAJOR order (b200_layernorm_fwd_windows /
 *      b200_window_rows); the cyclic shift and the window partition are that row order.  out: [B*H*W, C] bf16 = input of
 *      to_out, RASTER order.  lse: [B*H*W, heads] fp32 row log-sum-exp, window-major (nullable in inference). ------------- */
int b200_window_attn_fwd(const void* qkv, const float* pos_embedding, void* out, float* lse, int B, int H, int W, int C,
                         int heads, int shifted, void* stream);
int b200_window_attn_bwd_blocks(int B, int H, int W, int heads);
long long b200_window_attn_bwd_scratch_floats(int blocks); /* size (floats) of the `dpos_partial` scratch buffer */
/* backward from (qkv, lse, dout) alone: P is recomputed from the saved row log-sum-exp and rowsum(P o dP) stands in for
 * rowsum(dout o out), so the forward output is not an input.  dout: raster order; dqkv: window-major like qkv */
int b200_window_attn_bwd(const void* qkv, const float* pos_embedding, const float* lse, const void* dout,
                         void* dqkv, float* dpos, float* dpos_partial, int accumulate_dpos, int B, int H, int W, int C,
                         int heads, int shifted, void* stream);

/* ---- large-margin head (losses/large_margin.py:30-40 AddMarginProduct, :69-84 ArcMarginProduct) and the loss
 *      (losses/losses.py:22-28 FocalLoss; gamma = 0 == nn.CrossEntropyLoss mean) ---------------------------------- */
int b200_unit_rows(const float* x, void* out16, float* inv_norm, long long R, int E, long long ld_out, float eps, int as_f16,
                   void* stream); /* F.normalize rows -> bf16 (or fp16) */
int b200_margin_logits(const void* emb_unit, const void* w_unit, int B, int C, int E, const long long* label, float s, float m,
                       int kind /*0 arc, 1 cos*/, int easy_margin, float* logits, long long ldo, float* cos_label, void* stream);
int b200_margin_ce(const float* logits, long long ldl, const long long* label, const float* cos_label, int B, int C, float s,
                   float m, int kind, int easy_margin, float gamma, float* loss_rows, float* loss_mean, void* G, long long ldg,
                   float* rdot, float* cdot, void* stream);
/* stand-alone FocalLoss.forward(input, target) (losses/losses.py:22-28): loss_rows[b] = (1 - p_b)^gamma * nll_b, loss_mean = their
 * mean; dlogits (optional, fp32 [B, ldd]) = d loss_mean / d logits.  A label outside [0, C) yields a NaN row (no device sync). */
int b200_focal_loss(const float* logits, long long ldl, const long long* label, int B, int C, float gamma, float* loss_rows,
                    float* loss_mean, float* dlogits, long long ldd, void* stream);
int b200_unit_rows_bwd(const float* T, const float* x, const float* inv_norm, const float* dot, const float* scale_dev,
                       float* out_f32, void* out_bf16, long long R, int E, int accumulate, void* stream);

/* ---- fused multi-tensor optimizer step: torch.optim.SGD(momentum) as built at
 *      configs/dog_fe/fe_dogs_config.py:123-133, torch.optim.AdamW of configs/dog_fe/body_dog_fe.py:123-131 -------- */
int b200_opt_chunk_elems(void);
/* step_offset: optimizer steps taken since the device table was written (added to every B200OptTensor.step), so the table
 * is uploaded once and not once per step */
int b200_optimizer_step(int kind, const void* tensors_dev, const void* chunks_dev, int n_chunks, float grad_scale, int step_offset,
                        void* stream);
/* The same step with the gradient of every tensor taken as the sum of n_src copies: copy r of element i is read at
 * grad[src_shift + r * src_stride + i] (elements).  This is the data-parallel step: the copies are the ranks' slots of the
 * peer arena below, summed in slot order on every rank, so the all-reduce is folded into the optimizer pass. */
int b200_optimizer_step_sum(int kind, const void* tensors_dev, const void* chunks_dev, int n_chunks, float grad_scale, int step_offset,
                            int n_src, long long src_stride, long long src_shift, void* stream);

/* ---- gradient exchange over NVLink peer memory: replaces the bucketed NCCL all-reduce that Lightning's DDPPlugin /
 *      torch DDP run for the reference (utils/__init__.py:114-119, get_strategy).  One process per GPU; the arena of a rank is
 *      cudaMalloc'ed memory exported through CUDA IPC (64-byte handle), opened by the other ranks of the node, and written by
 *      them with asynchronous peer copies (copy engines, no SMs).  Host protocol: b200/peer.py. -------------------------- */
int b200_peer_alloc(long long bytes, void** ptr, unsigned char* handle64);  /* cudaMalloc + cudaIpcGetMemHandle            */
int b200_peer_open(const unsigned char* handle64, void** ptr);              /* cudaIpcOpenMemHandle (lazy peer access)     */
int b200_peer_close(void* ptr);                                             /* of a pointer returned by b200_peer_open     */
int b200_peer_free(void* ptr);                                              /* of a pointer returned by b200_peer_alloc    */
int b200_peer_copy(void* dst, const void* src, long long bytes, void* stream); /* cudaMemcpyAsync, either side may be a peer's */

/* ---- whole Swin backbone (models/swin.py:196-225): plan = shapes + buffer layout, no memory of its own --------- */
void* b200_swin_create(int batch, int img, int channels, int hidden_dim, const int* layers, const int* heads,
                       const int* downscaling, int num_classes, int head_dim, int window_size, int training);
void b200_swin_destroy(void* plan);
long long b200_swin_param_elems(const void* plan);
int b200_swin_param_count(const void* plan);
int b200_swin_param_offsets(const void* plan, long long* offsets, long long* numels, int n);
long long b200_swin_wcache_bytes(const void* plan);
long long b200_swin_workspace_bytes(const void* plan);
int b200_swin_sync_weights(const void* plan, const float* params, void* wcache, void* stream);
int b200_swin_forward(const void* plan, const float* params, const void* wcache, const void* img_nchw, int img_is_u8, float* emb,
                      void* workspace, long long workspace_bytes, void* stream); /* img: fp32 in [0,1], or uint8 (x/255 fused) */
int b200_swin_backward(const void* plan, const float* params, const void* wcache, const float* demb, float* grads,
                       void* workspace, long long workspace_bytes, int stage_hi, int stage_lo, void* stream);

/* ---- gallery matching: cosine scores + top-k (engine/controller.py:77-91, similarity_f of
 *      configs/dog_fe/fe_dogs_config.py:89-93; query != gallery form: generate_tsv_to_reproduce2.py:63-119) ---------- */
int b200_gallery_prepare(const float* emb, void* unit_f16, double* norm, long long n, int dim, void* stream);
/* Rows for the tensor-core pass in a frame fitted to a (concentrated) gallery, with error accounting for the certificate.
 * frame = [mu (dim) | w (dim) | |mu|] from b200_gallery_frame(mean unit gallery row): H = I - 2 w w^T maps mu / |mu| onto the
 * first axis and   gallery (role 1): row = scale * H (g^ - mu)      query (role 0): row = H q^,
 * so row_q . row_g / scale = cos(q, g) - q^ . mu: the ranking of the cosine, with operands of the size of the embeddings'
 * spread.  frame null: plain unit rows.  err (nullable, [n][4]) = {|d_1|, |d_rest|, |row_1|, |row_rest|} (d = fp16 rounding
 * residual, unscaled); stats (nullable, 4 floats the caller zeroes): maxima over the rows of {|row_1|, |row_rest|, |d_1|, |d_rest|}. */
int b200_gallery_prepare_ex(const float* emb, const float* frame, int role, float scale, void* f16_rows, double* norm, float* err,
                            float* stats, long long n, int dim, void* stream);
int b200_gallery_frame(const float* mean_unit_row, int dim, float* frame /* [2 * dim + 1] */, void* stream);
int b200_unit_row_mean_blocks(long long n);
int b200_unit_row_mean(const float* emb, long long n, int dim, float* mean, float* partial, void* stream); /* mean unit row */
long long b200_cosine_topk_workspace_bytes(long long nq, long long ng, int dim, int k);
int b200_cosine_topk(const float* q, const void* q_unit_f16, const double* q_norm, long long nq, const float* g,
                     const void* g_unit_f16, const double* g_norm, long long ng, int dim, int k, long long exclude_self_offset,
                     long long g_index_base, int* out_idx, double* out_score, void* workspace, long long workspace_bytes,
                     void* stream);
/* b200_cosine_topk + exactness certificate: a query whose kept candidates cannot be PROVEN to contain the true top-k (bound on
 * |fp16 score - exact score| from the measured rounding residuals vs the margin between the pruning thresholds and the exact
 * k-th score) is first re-ranked again with 512 instead of 128 exactly re-scored candidates and, if still unproven, re-done by
 * an exact fp64 scan of the gallery.  uncertified: int [1 + nq] = count of the queries that needed that scan, then their indices. */
int b200_cosine_topk_certified(const float* q, const void* q_f16, const double* q_norm, const float* q_err, long long nq,
                               const float* g, const void* g_f16, const double* g_norm, const float* frame, float g_scale,
                               const float* g_stats, long long ng, int dim, int k, long long exclude_self_offset, long long g_index_base,
                               int* out_idx, double* out_score, int* uncertified, void* workspace, long long workspace_bytes,
                               void* stream);
int b200_topk_merge(const double* scores, const int* idx, long long nq, int lists, int k_in, int k_out, int* out_idx,
                    double* out_score, void* stream);
/* Pair verification scores, engine/controller.py:60-68 + similarity_f (configs/dog_fe/fe_dogs_config.py:89-93):
 * out[p] = (cosine(emb[i1[p]], emb[i2[p]]) + 1) / 2 in fp32.  emb [n, dim] fp32, i1 / i2 int64 [n_pairs] (row numbers < n). */
int b200_pair_similarity(const float* emb, long long n, int dim, const long long* i1, const long long* i2, long long n_pairs,
                         float* out, void* stream);
int b200_recall_hits(const int* top_idx, long long nq, int k_stride, const long long* q_class, const long long* g_class,
                     const int* ks, int n_ks, unsigned long long* hits, void* stream);

/* ---- convolutional FE backbone: torchvision.models.resnet50 with fc -> 512, the model() of the reference's FE configs
 *      (configs/dog_fe/fe_dogs_config.py:96-109).  Activations: bf16 NHWC rows on a padded grid, (H + 2) x (W + 2) rows per
 *      image with a ring of zero rows (H = 0: plain [rows, C] matrix); see csrc/convnet.cu ------------------------------ */
/* Implicit convolution: out[r, n] = sum_t sum_c a[r + tap_shift[t], c] * b[n, t * C + c] (+ aux[r, n] with B200_EPI_RESID);
 * a [M, C] bf16 (C % 64 == 0), b [N, taps * C] bf16, out / aux [M, N] bf16; rows outside [0, M) read as zero.  A 3x3 convolution
 * on a padded grid is taps = 9, tap_shift = dy * (W + 2) + dx; its data gradient the same call with negated shifts. */
int b200_gemm_taps(const void* a, long long lda, const void* b, long long ldb, int M, int N, int C, int taps, const int* tap_shift,
                   int mode, void* out, long long ldo, const void* aux, long long ldaux, void* stream);
/* its weight gradient: partial [splits][N][taps * C], dW[n, t * C + c] = sum_p dy[p, n] * x[p + tap_shift[t], c]; splits as
 * b200_gemm_splits(tokens, splits) reports */
int b200_gemm_wgrad_taps(const void* dy, long long ldy, const void* x, long long ldx, long long tokens, int N, int C, int taps,
                         const int* tap_shift, float* partial, int splits, void* stream);
/* eval-mode convolution + BatchNorm (+ residual) (+ ReLU) in one launch: b = weights with the BatchNorm scale folded in, bias = its
 * shift; out[r, n] = 0 on ring rows of the (H, W) grid (H = 0: none), else act(a . b + bias[n] (+ aux[r, n])).  taps <= 1: 1x1. */
int b200_gemm_conv_bn(const void* a, long long lda, const void* b, long long ldb, int M, int N, int C, int taps, const int* tap_shift,
                      const float* bias, int relu, int H, int W, const void* aux, long long ldaux, void* out, long long ldo, void* stream);
int b200_bn_stats_blocks(long long rows);
/* nn.BatchNorm2d, training mode: batch statistics over the interior rows (count of them given) -> out [4][C] = scale, shift,
 * mean, rstd; running statistics (nullable) updated with `momentum` (unbiased variance).  scratch [blocks][2][C] floats. */
int b200_bn_stats(const void* x, long long rows, int C, int H, int W, double count, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float momentum, float eps, float* out, float* scratch, void* stream);
/* y = [relu](x * scale + shift [+ residual]) on interior rows, zero on the ring */
int b200_bn_apply(const void* x, const float* scale, const float* shift, const void* residual, int relu, long long rows, int C, int H,
                  int W, void* y, void* stream);
/* BatchNorm (+ ReLU when y is given; relu_from_x: the mask of y = relu(x * scale + shift) is recomputed from x) backward: sums [2][C] = dbeta, dgamma; dx; dz_out (nullable) = dy * [y > 0];
 * count = 0: frozen (eval-mode) statistics, dx = gamma * rstd * dz */
int b200_bn_backward(const void* dy, const void* y, int relu_from_x, const void* x, const float* stats, const float* gamma, long long rows,
                     int C, int H, int W, double count, void* dx, void* dz_out, float* sums, float* scratch, void* stream);
/* conv1 (7x7, stride 2, pad 3) patches: img [B, 3, H, W] uint8 (x / 255) or fp32 -> cols [B * H/2 * W/2, 160] bf16, column
 * (r * 7 + s) * 3 + c, columns 147..159 zero */
int b200_stem_im2col(const void* img, int is_u8, int B, int H, int W, void* cols, void* stream);
/* bn1 + relu + MaxPool2d(3, 2, 1) fused: a [B * IH * IW, C] -> y on the padded (IH / 2, IW / 2) grid, tap [B * IH/2 * IW/2, C] u8 */
int b200_stem_pool_fwd(const void* a, const float* scale, const float* shift, int B, int IH, int IW, int C, void* y, void* tap, void* stream);
int b200_stem_pool_bwd(const void* dy, const void* tap, const void* a, const float* scale, const float* shift, int B, int IH, int IW,
                       int C, void* dz, void* stream);
/* stride-2 sampling between padded grids (pixels (2 i, 2 j) of the (H, W) grid) and its adjoint */
int b200_grid_sample2(const void* src, int B, int H, int W, int C, int down, void* dst, void* stream);
/* patches of a stride-2 3x3 convolution: x on the (H, W) grid -> cols [rows of the (H / 2, W / 2) grid, 9 * C] (tap-major) */
int b200_grid_patches_s2(const void* x, int B, int H, int W, int C, void* cols, void* stream);
/* AdaptiveAvgPool2d(1) over the interior: forward x (grid) -> out [B, C]; backward x = d_out [B, C] -> out (grid) */
int b200_grid_avgpool(const void* x, int B, int H, int W, int C, int backward, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_FE_H_ */
