// tcgen05 GEMM core for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (16-bit operands, fp32 accumulation in TMEM),
// pluggable epilogue.
//
//   warp 0      TMA producer   : cp.async.bulk.tensor (SWIZZLE_128B boxes of 64 K-elements) -> smem ring (3-8 stages,
//                                sized at run time from block_n)
//   warp 1      MMA issuer     : one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=block_n, K=16)
//   warps 2..9  epilogue       : two warps per TMEM lane quarter, each taking a share of the tile's column boxes:
//                                [TMA-prefetched aux box (residual / saved GELU') ->]       tcgen05.ld -> registers ->
//                                Epi::compute (bias / GELU / residual / GELU' / margin ...) -> swizzled smem box
//                                (32 rows x <=128 B) -> TMA bulk tensor STORE (coalesced, asynchronous, clipped at the
//                                matrix edge); two boxes per warp, so loads / math / stores of consecutive boxes overlap
//
// Persistent: grid = min(#tiles, #SMs); every role walks the same static tile sequence
// (tile = blockIdx.x + i * gridDim.x; n fastest so the n-tiles of one A row-block run in the same wave
// and share it through L2).  Two TMEM accumulator stages (2 x 256 columns) let the MMA of tile i+1
// overlap the epilogue of tile i.  block_n is a RUN-TIME multiple of 16 (<= 256): it only enters through
// the B tensor map's box, the instruction descriptor and loop bounds, so one instantiation per epilogue
// serves every layer shape of the network.  Optional split-K (each split writes its own fp32 partial)
// serves the weight-gradient GEMMs whose contraction runs over all tokens.
#pragma once
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace gemm {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // 64 x 2 B = one 128-B swizzle row
constexpr int kMaxStages = 8;
constexpr int kMaxBlockN = 256;
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
constexpr int kMinPipeBytes = 3 * (kABytes + kMaxBlockN * kBlockK * 2);   // the operand ring always holds 3 stages at block_n = 256 (144 KB)
constexpr int kTmemCols = 512;                    // 2 accumulator stages x 256 fp32 columns
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;     // 320
constexpr int kBoxBytes = 32 * 128;               // one output box: 32 rows x (at most) 128 B
constexpr int kStagingPerWarp = 2 * kBoxBytes;    // double buffer (or the two outputs of the GELU epilogue)
constexpr int kBarBytes = 512;
constexpr int kOnesBytes = 64 * 128;                // the all-ones operand tile of the fused bias gradient (shares the bias staging area)
constexpr int kOnesCol = 240;                     // TMEM columns [240, 256) of an accumulator stage hold its product
constexpr int kBiasFloats = 4096;                 // the (zero-padded) bias vector is staged in smem once per CTA when it fits
// Every launch asks for the SM's whole 227 KB.  What the epilogue does not need - staging boxes, barriers and the ACTUAL
// size of the staged bias vector (or the all-ones tile) - goes to the operand ring, whose depth decides how many tiles of A
// are in flight: with 1.5 tiles (3 stages of a 192-wide tile at K = 96) the issue of a tile's last K block waits for the
// previous tile's MMA, the tile period equals the HBM latency and the epilogue warps starve on the accumulator barrier
// (ncu r01t, qkv of stage 1: 16 % of all samples); the fourth stage halves that period.
constexpr int kSmemBytes = 232448;
constexpr int kFixedSmem = kEpiWarps * kStagingPerWarp + kBarBytes + 1024 /*alignment slack*/;
static_assert(kMinPipeBytes + kFixedSmem + kBiasFloats * 4 <= kSmemBytes, "shared memory budget");

struct CoreParams {
  int M, N;                  // output rows (rows of A) / columns (rows of B)
  int block_n;               // multiple of 16, <= 256
  int m_blocks, n_blocks;
  int k_blocks;              // ceil(K / 64) over the whole contraction
  int splits;                // split-K factor (>= 1)
  int k_blocks_per_split;
  uint32_t idesc;            // tcgen05 instruction descriptor (formats, majors, M=128, N=block_n)
  uint32_t idesc_last;       // the same for the last (ragged) column block: N = columns left, rounded up to 16
  int k_last_steps;          // K=16 MMA steps of the last 64-wide K block (1..4): padding is not multiplied
  int mn_major;              // 1: operands are [K, M] / [K, N] row-major (contraction index = row): weight gradients
  int b_chunks;              // mn_major: 64-column chunks of the B tile = ceil(block_n / 64)
  int stages;                // operand ring depth
  int stage_bytes;           // 16 KB (A) + B tile, 1024-aligned
  int has_aux;               // epilogue consumes a bf16 [M, N] side input, streamed by TMA into the staging boxes
  int out_bytes;             // bytes per output element (2 = bf16, 4 = fp32)
  int box_cols;              // columns per output box (divides block_n; box row = box_cols * out_bytes in {32, 64, 128} B)
  int n_out;                 // 1, or 2 when the epilogue also emits a second tensor (GELU derivative)
  FastDiv div_n, div_m;      // tile -> (n block, m block, split)
  float* colsum_out;         // weight gradients only: [splits][M] column sums of the A operand (= the bias gradient), or null
  uint32_t idesc_ones;       // N = 16 instruction descriptor of the all-ones MMA that produces them
  int bias_smem;             // 1: the epilogue reads the bias from its smem copy (padded N <= kBiasFloats)
  int pipe_bytes;            // operand ring size = stages * stage_bytes (a multiple of 1024)
  uint32_t wait_ns;          // suspend-time hint of the producer / MMA-issuer barrier waits
  int m_pairs;               // CTA-pair variant: pairs of row blocks = ceil(m_blocks / 2)
  FastDiv div_mp;
  uint32_t idesc2, idesc2_last;   // M = 256 instruction descriptors of the cta_group::2 MMA
  // Implicit convolution (K-major only): the contraction runs over `taps` filter taps x kb_per_tap K blocks of channels; tap t
  // reads the rows of A shifted by tap_shift[t] (rows outside [0, M) are zero-filled by the TMA) - see b200_gemm_taps
  // MN-major (weight gradient of such a convolution): the OUTPUT columns are (tap, channel) pairs, tap_cols channels per tap;
  // a 64-column chunk of tap t reads the rows of B shifted by tap_shift[t] (one tile may span several taps: they share the A tile)
  int taps, kb_per_tap, tap_cols;
  int tap_shift[9];
};

// instruction descriptor, kind::f16: [4,6) D fmt (1=f32), [7,10) A fmt, [10,13) B fmt (0=f16, 1=bf16),
// [15] A major, [16] B major (0 = K-major), [17,23) N>>3, [24,29) M>>4
inline uint32_t make_idesc(bool is_bf16, int block_n, bool mn_major = false, int m = kBlockM) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (is_bf16 ? 1u : 0u) << 7;
  d |= (is_bf16 ? 1u : 0u) << 10;
  d |= (mn_major ? 1u : 0u) << 15;
  d |= (mn_major ? 1u : 0u) << 16;
  d |= static_cast<uint32_t>(block_n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}

__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// 8-bit storage of the GELU derivative: gelu'(x) lies in [-0.1290, 1.1290]; q = rint((g - kQ8Lo) * kQ8Scale) in 0..255 has an
// absolute error <= 0.0025 everywhere - the size of the bf16 rounding error for g near 1, where most of the mass is.
constexpr float kQ8Lo = -0.13f;
constexpr float kQ8Scale = 255.0f / 1.26f;
constexpr float kQ8Step = 1.26f / 255.0f;
__device__ __forceinline__ uint32_t q8_quant(float g) {
  uint32_t q;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q) : "f"(fmaf(g, kQ8Scale, -kQ8Lo * kQ8Scale)));
  return q;
}
__device__ __forceinline__ uint32_t q8_pack4(float a, float b, float c, float d) {
  const uint32_t lo = __byte_perm(q8_quant(a), q8_quant(b), 0x3340), hi = __byte_perm(q8_quant(c), q8_quant(d), 0x3340);
  return __byte_perm(lo, hi, 0x5410);
}
// byte k of w -> the float it encodes: 0x4B000000 | q is the float 2^23 + q; subtract, scale, shift
template <int K>
__device__ __forceinline__ float q8_dequant(uint32_t w) {
  const float m = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + K));
  return fmaf(m - 8388608.0f, kQ8Step, kQ8Lo);
}

// Compile-time specialisation of the epilogue (the small-K layers are bound by the epilogue's instruction issue rate):
//   Epi       what to compute per element          OUT_BYTES  2 = bf16 output, 4 = fp32 output
//   DUAL      second bf16 output (GELU derivative)       AUX   bf16 side input streamed by TMA into the staging boxes
//   Q8        the second output (DUAL) / the side input (AUX) is the GELU derivative quantised to 8 bits (see q8_*): a quarter
//             less traffic for the fc1 epilogue and 1.2 GB less saved activations at batch 256.  Measured (r01v): NOT faster -
//             fc1 1.69 -> 1.79 ms per step, the epilogue is bound by instruction issue / latency, not by the stores - so the
//             training plan keeps the bf16 derivative; the modes stay available (B200_EPI_GELU_Q8 / B200_EPI_DGELU_Q8)
//   EW        epilogue warps: 8 (two per scheduler, double-buffered staging boxes) or 16 (four per scheduler, one staging box
//             each, plain epilogues only).  The small-K layers are bound by the latency chain of a box (barrier -> tcgen05.ld
//             -> convert -> st.shared -> fence -> TMA store), and two warps per scheduler do not cover it.
//   CG2       CTA-pair variant (cta_group::2) for the compute-bound layers (K >= 384: stages 3-4).  A single-CTA MMA reads
//             (128 + N) x 64 x 2 B of shared memory per K block while the TMA writes the same amount: at N = 256 that is 192 B per
//             clock against the 128 B / clock the SM's shared memory delivers - the measured 47-65 % of tensor peak.  In a pair each
//             CTA keeps its own 128 rows of A but only HALF of the B tile (the tensor cores of both SMs see all of it), the leader
//             issues M = 256 MMAs for both, and the per-SM shared-memory traffic drops to the 128 B / clock that fits.
template <class Epi, int OUT_BYTES, bool DUAL, bool AUX, bool Q8 = false, int EW = kEpiWarps, bool CG2 = false>
__global__ void __launch_bounds__(64 + EW * 32, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_o2,
               const __grid_constant__ CUtensorMap tmap_aux, const CoreParams p, const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B wants 1024-B alignment
  const uint32_t staging_base = smem_base + static_cast<uint32_t>(p.pipe_bytes);
  const uint32_t bar_base = staging_base + kEpiWarps * kStagingPerWarp;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 2 + a); };
  auto aux_bar = [&](int w, int b) { return bar_base + 8u * (2 * kMaxStages + 4 + 2 * w + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4 + 2 * EW);
  constexpr int kStg = kEpiWarps * kStagingPerWarp / EW;      // staging bytes per epilogue warp: 8 KB (two boxes) or 4 KB (one)
  static_assert(EW == 8 || (EW == 16 && !DUAL && !AUX), "16 epilogue warps: plain epilogues only");
  const uint32_t bias_base = bar_base + kBarBytes;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  static_assert(!CG2 || EW == 8, "the CTA-pair variant runs eight epilogue warps");
  const uint32_t cta_rank = CG2 ? cluster_ctarank() : 0u;        // 0 = leader: issues the MMAs of the pair
  const bool leader = cta_rank == 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_o);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(full_bar(s), CG2 ? 2 : 1);   // pair: the leader's barrier counts both producers (and both CTAs' bytes)
      mbar_init(empty_bar(s), 1);
    }
    for (int w = 0; w < EW; ++w) {
      mbar_init(aux_bar(w, 0), 1);
      mbar_init(aux_bar(w, 1), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), CG2 ? 2 * EW : EW);   // one arrival per epilogue warp (pair: of both CTAs, on the leader's barrier)
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if (CG2) { tmem_alloc_cg2(tmem_slot, kTmemCols); tmem_relinquish_cg2(); }
    else { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  }
  if (p.colsum_out != nullptr) {
    // Bias gradient for free: db[m] = sum_t dY[t, m] is the weight-gradient GEMM with an all-ones second operand.  One
    // 8-KB tile of bf16 1.0 (layout and swizzle are irrelevant: every element is equal) sits where the bias staging would
    // be, and each K block gets four extra N = 16 MMAs into 16 spare TMEM columns of the accumulator stage.
    for (int i = threadIdx.x; i < kOnesBytes / 16; i += 64 + EW * 32) st_shared_v4(bias_base + 16u * i, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (CG2) cluster_sync_all();      // the peer's barriers are initialised before anything arrives on them from this CTA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_grid_sync();      // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail

  // The tile sequence: tile = tile0 + i * tstride.  Single CTA: one 128-row block per tile.  Pair: a "tile" is a pair of
  // row blocks (2 m_pair, 2 m_pair + 1) x one column block, walked by both CTAs of the cluster; CTA `rank` owns row block
  // 2 m_pair + rank (a row block past the end computes on zero-filled rows and its stores are clipped).
  const int num_tiles = CG2 ? p.m_pairs * p.n_blocks : p.m_blocks * p.n_blocks * p.splits;
  const int tile0 = CG2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tstride = CG2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  auto decode = [&](int tile, int& n_blk, int& m_blk, int& split) {
    const uint32_t t1 = p.div_n.div(static_cast<uint32_t>(tile));            // tile / n_blocks
    n_blk = tile - static_cast<int>(t1) * p.n_blocks;
    if (CG2) {
      m_blk = 2 * static_cast<int>(t1) + static_cast<int>(cta_rank);
      split = 0;
      return;
    }
    const uint32_t t2 = p.div_m.div(t1);                                     // tile / (n_blocks * m_blocks)
    m_blk = static_cast<int>(t1) - static_cast<int>(t2) * p.m_blocks;
    split = static_cast<int>(t2);
  };
  // K-major: one [128 x 64] A box + one [block_n x 64] B box per stage.  MN-major: the tile is [64 contraction rows x
  // 128 / block_n columns], loaded as [64 x 64] boxes (one 128-B swizzle row = 64 columns) 8 KB apart.
  constexpr uint32_t kChunk = 64 * 64 * 2;
  const uint32_t tx_bytes = p.mn_major ? static_cast<uint32_t>((2 + p.b_chunks) * kChunk)
                                       : static_cast<uint32_t>((kBlockM + p.block_n) * kBlockK * 2);       // pair: both CTAs' A + the whole B

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tstride) {
        int n_blk, m_blk, split;
        decode(tile, n_blk, m_blk, split);
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.k_blocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_backoff(empty_bar(stage), phase ^ 1u, p.wait_ns);
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          int a_col = kb * kBlockK, a_row = m_blk * kBlockM;
          if (p.taps > 0) {                       // implicit convolution: K block -> (filter tap, channel block)
            const int t = kb / p.kb_per_tap;
            a_col = (kb - t * p.kb_per_tap) * kBlockK;
            a_row += p.tap_shift[t];
          }
          if (CG2) {
            // this CTA's 128 rows of A and its half of the B tile; the bytes of both CTAs complete on the leader's barrier
            if (leader) mbar_arrive_expect_tx(full_bar(stage), 2u * kABytes + static_cast<uint32_t>(p.block_n * kBlockK * 2));
            tma_load_2d_cg2(sa, &tmap_a, full_bar(stage), a_col, a_row);
            tma_load_2d_cg2(sa + kABytes, &tmap_b, full_bar(stage), kb * kBlockK, n_blk * p.block_n + static_cast<int>(cta_rank) * (p.block_n >> 1));
            if (!leader) mbar_arrive_remote(full_bar(stage), 0u);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            continue;
          }
          mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
          if (!p.mn_major) {
            tma_load_2d(sa, &tmap_a, full_bar(stage), a_col, a_row);
            tma_load_2d(sa + kABytes, &tmap_b, full_bar(stage), kb * kBlockK, n_blk * p.block_n);
          } else {
            for (int c = 0; c < 2; ++c) tma_load_2d(sa + c * kChunk, &tmap_a, full_bar(stage), m_blk * kBlockM + c * 64, kb * kBlockK);
            for (int c = 0; c < p.b_chunks; ++c) {
              int b_col = n_blk * p.block_n + c * 64, b_row = kb * kBlockK;
              if (p.taps > 0) {                   // 64-column chunk -> (filter tap, channel chunk): shifted rows of the activation tensor
                const int t = b_col / p.tap_cols;
                b_col -= t * p.tap_cols;
                b_row += p.tap_shift[t < p.taps ? t : 0];
                if (t >= p.taps) b_col = p.tap_cols;        // past the last tap (ragged last tile): out of range -> zero fill
              }
              tma_load_2d(sa + kABytes + c * kChunk, &tmap_b, full_bar(stage), b_col, b_row);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && CG2) {
      if (leader) {
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = tile0; tile < num_tiles; tile += tstride) {
          mbar_wait_backoff(tempty_bar(acc), acc_phase ^ 1u, p.wait_ns);      // both CTAs' epilogues have drained this stage
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kMaxBlockN);
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait_backoff(full_bar(stage), phase, p.wait_ns);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * p.stage_bytes;
            const uint64_t da = make_sw128_desc(sa, 16, 1024);
            const uint64_t db = make_sw128_desc(sa + kABytes, 16, 1024);
            const int ksteps = (kb == p.k_blocks - 1) ? p.k_last_steps : kBlockK / 16;
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              if (k < ksteps) umma_f16_cg2(d_tmem, da + 2u * k, db + 2u * k, p.idesc2, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_cg2(empty_bar(stage));                   // the slot is free again in BOTH CTAs
            if (kb == p.k_blocks - 1) umma_commit_cg2(tfull_bar(acc));
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
        // the peer's last arrivals on this CTA's barriers must have landed before it may exit
        for (int a2 = 0; a2 < 2; ++a2) {
          mbar_wait_backoff(tempty_bar(acc), acc_phase ^ 1u, p.wait_ns);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    } else if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tstride) {
        int n_blk, m_blk, split;
        decode(tile, n_blk, m_blk, split);
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.k_blocks_per_split);
        const uint32_t idesc = (n_blk == p.n_blocks - 1) ? p.idesc_last : p.idesc;
        mbar_wait_backoff(tempty_bar(acc), acc_phase ^ 1u, p.wait_ns);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kMaxBlockN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_backoff(full_bar(stage), phase, p.wait_ns);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          // K-major : SBO = 1024 B between 8-row groups; a K=16 slice is +32 B inside the 128-B swizzle row (+2 in addr>>4)
          // MN-major: LBO = 8 KB between 64-column chunks, SBO = 1024 B between 8-row (contraction) groups; a K=16 slice
          //           is 16 rows = +2048 B (+128 in addr>>4)
          const uint64_t da = p.mn_major ? make_sw128_desc(sa, kChunk, 1024) : make_sw128_desc(sa, 16, 1024);
          const uint64_t db = p.mn_major ? make_sw128_desc(sa + kABytes, kChunk, 1024) : make_sw128_desc(sa + kABytes, 16, 1024);
          const uint32_t kstep = p.mn_major ? 128u : 2u;
          const int ksteps = (kb == p.k_blocks - 1) ? p.k_last_steps : kBlockK / 16;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            if (k < ksteps) umma_f16(d_tmem, da + kstep * k, db + kstep * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          if (p.colsum_out != nullptr && n_blk == 0) {
            const uint64_t d1 = make_sw128_desc(bias_base, kChunk, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              if (k < ksteps) umma_f16(d_tmem + kOnesCol, da + kstep * k, d1 + 128u * k, p.idesc_ones, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));                 // smem slot reusable once these MMAs retire
          if (kb == kb1 - 1) umma_commit(tfull_bar(acc)); // accumulator complete
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (kb1 <= kb0) umma_commit(tfull_bar(acc));      // degenerate split: nothing to add
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    const int ew = warp - 2;                // epilogue warp 0..7
    const int q = warp & 3;                 // TMEM lane quarter this warp may access (hardware rule: warp_id % 4)
    const int nboxes = p.block_n / p.box_cols;
    constexpr int kWpq = EW / 4;                                  // the warps of a lane quarter split the column boxes
    const int wq = ew >> 2;
    const int box_lo = (nboxes * wq + kWpq - 1) / kWpq;
    const int box_hi = (nboxes * (wq + 1) + kWpq - 1) / kWpq;
    const uint32_t stg = staging_base + ew * kStg;
    const uint32_t row_bytes = static_cast<uint32_t>(p.box_cols * OUT_BYTES);
    const uint32_t box_bytes = 32u * row_bytes;
    // this lane's row inside a staging box, and its swizzle term: the 16-B chunk index is XORed with address bits 7.. of
    // the row (Swizzle<3|2|1,4,3> of the 128 / 64 / 32-B TMA modes); both are loop invariants
    const uint32_t row_off = static_cast<uint32_t>(lane) * row_bytes;
    const uint32_t swz = ((row_off >> 7) & ((row_bytes >> 4) - 1u)) << 4;
    // the 8-bit tensor (Q8): one byte per element, so its box rows are box_cols bytes with their own swizzle term
    const uint32_t row_off8 = static_cast<uint32_t>(lane * p.box_cols);
    const uint32_t swz8 = ((row_off8 >> 7) & ((static_cast<uint32_t>(p.box_cols) >> 4) - 1u)) << 4;
    constexpr int kChunkBytes = 16 * OUT_BYTES;                   // bytes one 16-column chunk occupies in a box row
    constexpr int kMaxChunks = 128 / kChunkBytes;                 // 4 (bf16) / 2 (fp32)
    const bool aux = AUX && box_hi > box_lo;
    int buf = 0;
    uint32_t aux_phase0 = 0, aux_phase1 = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    // stream the aux box of (tile, box) into staging buffer b (same geometry / swizzle as the output box)
    auto issue_aux = [&](int tile, int box, int b) {
      int n_blk, m_blk, split_unused;
      decode(tile, n_blk, m_blk, split_unused);
      mbar_arrive_expect_tx(aux_bar(ew, b), Q8 ? 32u * static_cast<uint32_t>(p.box_cols) : box_bytes);
      tma_load_3d(stg + static_cast<uint32_t>(b) * kBoxBytes, &tmap_aux, aux_bar(ew, b), n_blk * p.block_n + box * p.box_cols,
                  m_blk * kBlockM + q * 32, 0);
    };
    if (aux && lane == 0 && tile0 < num_tiles) issue_aux(tile0, box_lo, 0);
    // the epilogue's per-column vector (bias), zero-padded to the tile grid, once per CTA: the epilogue then reads it with
    // broadcast shared loads and needs no column guard
    const float* bias_s = reinterpret_cast<const float*>(smem_raw + (bias_base - smem_u32(smem_raw)));
    if (p.bias_smem) {
      Epi::stage_columns(ep, p, const_cast<float*>(bias_s), static_cast<int>(threadIdx.x) - 64, EW * 32);
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
    }
    for (int tile = tile0; tile < num_tiles; tile += tstride) {
      int n_blk, m_blk, split;
      decode(tile, n_blk, m_blk, split);
      const int kb0 = split * p.k_blocks_per_split;
      const bool has_k = min(p.k_blocks, kb0 + p.k_blocks_per_split) > kb0;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * kBlockM + q * 32 + lane;
      const int col_tile = n_blk * p.block_n;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * kMaxBlockN);
      for (int box = box_lo; box < box_hi; ++box) {
        const int c_tile = box * p.box_cols;                     // first column of the box inside the tile
        if (!AUX && col_tile + c_tile >= p.N) continue;          // padded tile: nothing of this box is inside the matrix
        if (lane == 0) {
          if (AUX) {
            // the OTHER buffer was last read by the store of the previous box: drain it, then stream the next aux box in
            bulk_wait_read<0>();
            if (aux) {
              const bool more_here = box + 1 < box_hi;
              const int ntile = more_here ? tile : tile + tstride;
              if (ntile < num_tiles) issue_aux(ntile, more_here ? box + 1 : box_lo, buf ^ 1);
            }
          } else if (DUAL) {
            bulk_wait_read<0>();               // both buffers are rewritten for every box
          } else if (EW == 16) {
            bulk_wait_read<0>();               // one staging box per warp: the other warps of the scheduler cover the wait
          } else {
            bulk_wait_read<1>();               // the buffer about to be rewritten was read by the store before last
          }
        }
        __syncwarp();
        // DUAL: one 128-B-wide box per output, no double buffering (measured, r01p: halving the boxes to double-buffer them
        // costs more in per-box overhead than the exposed store latency it hides: fc1 305 -> 343 us)
        const uint32_t sbuf = stg + ((DUAL || EW == 16) ? 0u : static_cast<uint32_t>(buf) * kBoxBytes);
        if (AUX) {
          if (buf == 0) { mbar_wait(aux_bar(ew, 0), aux_phase0); aux_phase0 ^= 1u; }
          else { mbar_wait(aux_bar(ew, 1), aux_phase1); aux_phase1 ^= 1u; }
        }
        const uint32_t srow = sbuf + row_off;
        uint32_t aux8[kMaxChunks][4];
        if (AUX && Q8) {
          // the 1-byte side input shares the buffer the 2-byte result is about to overwrite, with a different row pitch:
          // every lane pulls its whole row into registers before anybody writes
#pragma unroll
          for (int ci = 0; ci < kMaxChunks; ++ci)
            if (ci * 16 < p.box_cols) ld_shared_v4(sbuf + row_off8 + ((ci * 16) ^ swz8), aux8[ci][0], aux8[ci][1], aux8[ci][2], aux8[ci][3]);
          __syncwarp();
        }
        // all TMEM loads of the box are issued before the first wait: their latencies overlap instead of adding up
        // (with 16 epilogue warps the register budget is 112 per thread: one chunk at a time, the other warps overlap)
        uint32_t r[kMaxChunks][16];
        if (EW != 16) {
#pragma unroll
          for (int ci = 0; ci < kMaxChunks; ++ci)
            if (ci * 16 < p.box_cols) tmem_ld16(taddr + c_tile + ci * 16, r[ci]);
          tmem_ld_wait();
        }
#pragma unroll
        for (int ci = 0; ci < kMaxChunks; ++ci) {
          if (ci * 16 < p.box_cols) {
            if (EW == 16) {
              tmem_ld16(taddr + c_tile + ci * 16, r[ci]);
              tmem_ld_wait();
            }
            float ax[16];
            if (AUX && Q8) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                ax[4 * k + 0] = q8_dequant<0>(aux8[ci][k]); ax[4 * k + 1] = q8_dequant<1>(aux8[ci][k]);
                ax[4 * k + 2] = q8_dequant<2>(aux8[ci][k]); ax[4 * k + 3] = q8_dequant<3>(aux8[ci][k]);
              }
            } else if (AUX) {      // the bf16 side input sits exactly where the result will be written
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                uint32_t w0, w1, w2, w3;
                ld_shared_v4(srow + ((ci * kChunkBytes + 16 * k) ^ swz), w0, w1, w2, w3);
                ax[8 * k + 0] = bf16lo(w0); ax[8 * k + 1] = bf16hi(w0); ax[8 * k + 2] = bf16lo(w1); ax[8 * k + 3] = bf16hi(w1);
                ax[8 * k + 4] = bf16lo(w2); ax[8 * k + 5] = bf16hi(w2); ax[8 * k + 6] = bf16lo(w3); ax[8 * k + 7] = bf16hi(w3);
              }
            }
            if (!has_k) {
#pragma unroll
              for (int i = 0; i < 16; ++i) r[ci][i] = 0u;
            }
            float o[16], o2[16];
            Epi::template compute<DUAL>(ep, p, bias_s, row, col_tile + c_tile + ci * 16, reinterpret_cast<const float(&)[16]>(r[ci]), ax, o, o2);
            if (OUT_BYTES == 2) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const uint32_t sw = srow + ((ci * kChunkBytes + 16 * k) ^ swz);
                st_shared_v4(sw, pack_bf16(o[8 * k], o[8 * k + 1]), pack_bf16(o[8 * k + 2], o[8 * k + 3]),
                             pack_bf16(o[8 * k + 4], o[8 * k + 5]), pack_bf16(o[8 * k + 6], o[8 * k + 7]));
                if (DUAL && !Q8)
                  st_shared_v4(sw + kBoxBytes, pack_bf16(o2[8 * k], o2[8 * k + 1]), pack_bf16(o2[8 * k + 2], o2[8 * k + 3]),
                               pack_bf16(o2[8 * k + 4], o2[8 * k + 5]), pack_bf16(o2[8 * k + 6], o2[8 * k + 7]));
              }
              if (DUAL && Q8)
                st_shared_v4(sbuf + kBoxBytes + row_off8 + ((ci * 16) ^ swz8), q8_pack4(o2[0], o2[1], o2[2], o2[3]), q8_pack4(o2[4], o2[5], o2[6], o2[7]),
                             q8_pack4(o2[8], o2[9], o2[10], o2[11]), q8_pack4(o2[12], o2[13], o2[14], o2[15]));
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                st_shared_v4(srow + ((ci * kChunkBytes + 16 * k) ^ swz), __float_as_uint(o[4 * k]), __float_as_uint(o[4 * k + 1]),
                             __float_as_uint(o[4 * k + 2]), __float_as_uint(o[4 * k + 3]));
            }
          }
        }
        fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          const int gc = col_tile + c_tile, gr = m_blk * kBlockM + q * 32;
          tma_store_3d(&tmap_o, sbuf, gc, gr, split);
          if (DUAL) tma_store_3d(&tmap_o2, sbuf + kBoxBytes, gc, gr, split);
          bulk_commit();
        }
        buf ^= 1;
      }
      if (p.colsum_out != nullptr && n_blk == 0 && ew < 4) {       // one warp per lane quarter: column 0 of the ones product
        uint32_t r[16];
        tmem_ld16(taddr + kOnesCol, r);
        tmem_ld_wait();
        if (row < p.M) p.colsum_out[1LL * split * p.M + row] = has_k ? __uint_as_float(r[0]) : 0.0f;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG2 && !leader) mbar_arrive_remote(tempty_bar(acc), 0u);     // the leader's MMA thread waits for both CTAs
        else mbar_arrive(tempty_bar(acc));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) bulk_wait_all();          // smem must stay valid (and the writes must land) before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (CG2) cluster_sync_all();               // no CTA of the pair leaves (or frees TMEM) while the other can still reach it
  if (warp == 2) {
    tc_fence_after();
    if (CG2) tmem_dealloc_cg2(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int encode_tmap_2d(CUtensorMap* map, bool is_bf16, const void* ptr, uint64_t inner, uint64_t outer,
                   uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer);
// output tensor [splits][rows][cols] (elem_bytes 2 = bf16, 4 = fp32), box = [box_cols x 32 rows x 1], swizzle = box row bytes
int encode_tmap_out(CUtensorMap* map, int elem_bytes, const void* ptr, uint64_t cols, uint64_t rows, uint64_t splits,
                    uint64_t pitch_elems, uint64_t split_stride_elems, uint32_t box_cols);

int pick_block_n(int N, int K);

struct Operands {
  const void* a; int lda;     // [M, K] row-major, 16-bit           (mn_major: [K, M] row-major)
  const void* b; int ldb;     // [N, K] row-major, 16-bit           (mn_major: [K, N] row-major)
  int M, N, K;
  bool is_bf16;
  int block_n;                // 0 = auto
  int splits;                 // <= 1 = no split-K
  int max_ctas;               // 0 = #SMs
  bool mn_major = false;      // D = A^T B with the contraction over the ROWS of both operands (weight gradients)
  int taps = 0;               // > 0: implicit convolution - A is [M, K / taps], tap t contributes its rows shifted by tap_shift[t]
  int tap_shift[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

struct Output {
  void* ptr; long long ld; int elem_bytes;          // primary output [M, N] (per split: + split_stride elements)
  void* ptr2; long long ld2;                         // optional second bf16 output (GELU derivative), or nullptr
  long long split_stride;                            // elements between split partials (0 when splits == 1)
  const void* aux; long long ldaux;                 // optional bf16 [M, N] epilogue input (residual / saved GELU derivative)
  float* colsum = nullptr;                           // weight gradients: [splits][M] column sums of A (fused bias gradient)
};

inline int pick_box_cols(int block_n, int elem_bytes) {
  const int widest = 128 / elem_bytes;               // 64 bf16 / 32 fp32 columns = one 128-B row
  static const int forced = [] { const char* e = getenv("B200_BOX_COLS"); return e != nullptr ? atoi(e) : 0; }();   // experiments
  if (forced > 0 && forced <= widest && block_n % forced == 0 && forced * elem_bytes >= 32) return forced;
  for (int w = widest; w >= 32; w >>= 1)             // prefer an even box count: the two warps of a quarter split it evenly
    if (block_n % w == 0 && ((block_n / w) & 1) == 0) return w;
  for (int w = widest; w >= 16; w >>= 1)
    if (block_n % w == 0) return w;
  return 16;
}

int b200_cg2();   // 1 (default): compute-bound K-major layers run as CTA pairs; B200_CG2=0 -> single CTAs

template <class Epi, int OUT_BYTES, bool DUAL, bool AUX, bool Q8 = false, int EW = kEpiWarps, bool CG2 = false>
int launch(const Operands& o, const Output& out, const typename Epi::Params& ep, cudaStream_t stream) {
  static_assert(!Q8 || ((DUAL || AUX) && OUT_BYTES == 2), "Q8 qualifies the second output / the side input of a bf16 epilogue");
  B200_REQUIRE(out.elem_bytes == OUT_BYTES && (out.ptr2 != nullptr) == DUAL && (out.aux != nullptr) == AUX, "gemm: epilogue specialisation mismatch");
  B200_REQUIRE(o.M > 0 && o.N > 0 && o.K > 0, "gemm: empty problem M=%d N=%d K=%d", o.M, o.N, o.K);
  B200_REQUIRE(o.lda % 8 == 0 && o.ldb % 8 == 0, "gemm: row pitch must be a multiple of 8 elements (16 B)");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(o.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(o.b) & 15) == 0,
               "gemm: operand base must be 16-B aligned");
  CoreParams p;
  p.M = o.M;
  p.N = o.N;
  p.block_n = o.block_n > 0 ? o.block_n : pick_block_n(o.N, o.K);
  B200_REQUIRE(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= kMaxBlockN, "gemm: bad block_n %d", p.block_n);
  p.m_blocks = (o.M + kBlockM - 1) / kBlockM;
  p.n_blocks = (o.N + p.block_n - 1) / p.block_n;
  p.k_blocks = (o.K + kBlockK - 1) / kBlockK;
  p.splits = o.splits > 1 ? o.splits : 1;
  if (p.splits > p.k_blocks) p.splits = p.k_blocks;
  p.k_blocks_per_split = (p.k_blocks + p.splits - 1) / p.splits;
  p.splits = (p.k_blocks + p.k_blocks_per_split - 1) / p.k_blocks_per_split;   // no empty splits
  p.idesc = make_idesc(o.is_bf16, p.block_n, o.mn_major);
  const int n_last = o.N - (p.n_blocks - 1) * p.block_n;
  p.idesc_last = make_idesc(o.is_bf16, std::min(p.block_n, (n_last + 15) / 16 * 16), o.mn_major);
  p.k_last_steps = (o.K - (p.k_blocks - 1) * kBlockK + 15) / 16;
  p.mn_major = o.mn_major ? 1 : 0;
  p.b_chunks = (p.block_n + 63) / 64;
  const int b_bytes = o.mn_major ? p.b_chunks * 8192 : (CG2 ? p.block_n / 2 : p.block_n) * kBlockK * 2;
  p.stage_bytes = (kABytes + b_bytes + 1023) / 1024 * 1024;
  p.m_pairs = (p.m_blocks + 1) / 2;
  p.div_mp = make_fastdiv(static_cast<uint32_t>(p.m_pairs));
  p.idesc2 = make_idesc(o.is_bf16, p.block_n, false, 256);
  p.idesc2_last = p.idesc2;
  if (CG2)
    B200_REQUIRE(!o.mn_major && p.splits == 1 && o.N % p.block_n == 0 && p.block_n % 32 == 0 && out.colsum == nullptr,
                 "gemm: the CTA-pair variant needs a K-major, unsplit problem whose N is a multiple of the tile width");

  p.taps = o.taps;
  p.kb_per_tap = 0;
  p.tap_cols = 0;
  if (o.taps > 0) {
    B200_REQUIRE(o.taps <= 9, "gemm: at most 9 filter taps");
    if (!o.mn_major) {
      B200_REQUIRE(o.K % (o.taps * kBlockK) == 0 && p.splits == 1, "gemm: implicit convolution needs K / taps a multiple of 64 (K=%d taps=%d)", o.K, o.taps);
      p.kb_per_tap = o.K / o.taps / kBlockK;
    } else {
      p.tap_cols = o.N / o.taps;
      B200_REQUIRE(o.N % o.taps == 0 && p.tap_cols % 64 == 0 && p.block_n % 64 == 0, "gemm: tap weight gradient needs channels per tap (%d) and the tile width (%d) in multiples of 64",
                   p.tap_cols, p.block_n);
    }
    for (int t = 0; t < 9; ++t) p.tap_shift[t] = o.tap_shift[t];
  }
  CUtensorMap ta, tb;
  int rc;
  if (!o.mn_major) {
    rc = encode_tmap_2d(&ta, o.is_bf16, o.a, o.taps > 0 ? o.K / o.taps : o.K, o.M, o.lda, kBlockK, kBlockM);
    if (rc) return rc;
    rc = encode_tmap_2d(&tb, o.is_bf16, o.b, o.K, o.N, o.ldb, kBlockK, CG2 ? p.block_n / 2 : p.block_n);
  } else {
    rc = encode_tmap_2d(&ta, o.is_bf16, o.a, o.M, o.K, o.lda, 64, kBlockK);
    if (rc) return rc;
    rc = encode_tmap_2d(&tb, o.is_bf16, o.b, o.taps > 0 ? o.N / o.taps : o.N, o.K, o.ldb, 64, kBlockK);
  }
  if (rc) return rc;

  B200_REQUIRE(out.ptr != nullptr && (reinterpret_cast<uintptr_t>(out.ptr) & 15) == 0, "gemm: output must be 16-B aligned");
  B200_REQUIRE((out.ld * out.elem_bytes) % 16 == 0 && (out.split_stride * out.elem_bytes) % 16 == 0, "gemm: output pitch must be a multiple of 16 B");
  p.out_bytes = out.elem_bytes;
  p.box_cols = pick_box_cols(p.block_n, out.elem_bytes);
  if (Q8) B200_REQUIRE(p.box_cols >= 32, "gemm: the 8-bit GELU-derivative tensor needs boxes of >= 32 columns (block_n %d)", p.block_n);
  p.n_out = out.ptr2 != nullptr ? 2 : 1;
  p.has_aux = out.aux != nullptr ? 1 : 0;
  p.colsum_out = out.colsum;
  p.idesc_ones = make_idesc(o.is_bf16, 16, o.mn_major);
  if (out.colsum != nullptr)
    B200_REQUIRE(o.mn_major && o.is_bf16 && OUT_BYTES == 4 && p.block_n <= kOnesCol && !Epi::wants_columns(ep),
                 "gemm: fused column sums need a bf16 weight-gradient launch with block_n <= %d (got %d)", kOnesCol, p.block_n);
  p.div_n = make_fastdiv(static_cast<uint32_t>(p.n_blocks));
  p.div_m = make_fastdiv(static_cast<uint32_t>(p.m_blocks));
  p.bias_smem = (Epi::wants_columns(ep) && 1LL * p.n_blocks * p.block_n <= kBiasFloats) ? 1 : 0;
  const int extra_bytes = p.bias_smem ? p.n_blocks * p.block_n * 4 : (out.colsum != nullptr ? kOnesBytes : 0);
  p.stages = (kSmemBytes - kFixedSmem - (extra_bytes + 1023) / 1024 * 1024) / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  p.pipe_bytes = p.stages * p.stage_bytes;
  p.wait_ns = static_cast<uint32_t>(b200_wait_ns());
  CUtensorMap to, to2, tx;
  rc = encode_tmap_out(&to, out.elem_bytes, out.ptr, o.N, o.M, p.splits, out.ld, p.splits > 1 ? out.split_stride : 1LL * o.M * out.ld,
                       p.box_cols);
  if (rc) return rc;
  if (out.ptr2 != nullptr) {
    B200_REQUIRE(out.elem_bytes == 2 && (out.ld2 * (Q8 ? 1 : 2)) % 16 == 0 && (reinterpret_cast<uintptr_t>(out.ptr2) & 15) == 0, "gemm: second output must be 16-B aligned");
    rc = encode_tmap_out(&to2, Q8 ? 1 : 2, out.ptr2, o.N, o.M, 1, out.ld2, 1LL * o.M * out.ld2, p.box_cols);
    if (rc) return rc;
  } else {
    to2 = to;
  }
  tx = to;
  if (out.aux != nullptr) {
    B200_REQUIRE(out.elem_bytes == 2 && out.ptr2 == nullptr && p.splits == 1, "gemm: an aux input needs a single bf16 output");
    B200_REQUIRE((out.ldaux * (Q8 ? 1 : 2)) % 16 == 0 && (reinterpret_cast<uintptr_t>(out.aux) & 15) == 0, "gemm: aux must be 16-B aligned");
    rc = encode_tmap_out(&tx, Q8 ? 1 : 2, out.aux, o.N, o.M, 1, out.ldaux, 1LL * o.M * out.ldaux, p.box_cols);
    if (rc) return rc;
  }

  static bool attr_done = false;   // per instantiation
  if (!attr_done) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<Epi, OUT_BYTES, DUAL, AUX, Q8, EW, CG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done = true;
  }
  const long long tiles = CG2 ? 1LL * p.m_pairs * p.n_blocks : 1LL * p.m_blocks * p.n_blocks * p.splits;
  int ctas = o.max_ctas > 0 ? o.max_ctas : b200_num_sms();
  if (CG2) ctas /= 2;                       // clusters of two
  if (tiles < ctas) ctas = static_cast<int>(tiles);
  // A ragged last column block makes the tiles of one row block unequal (qkv of stage 1: 192 + 96 columns).  With the static
  // tile sequence (tile = cta + i * ctas, n fastest) a CTA count sharing a factor with n_blocks would hand some CTAs only
  // wide tiles and others only narrow ones; a coprime count rotates every CTA through all column blocks.
  if (p.n_blocks > 1 && o.N % p.block_n != 0 && tiles > ctas) {
    auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
    while (ctas > 1 && gcd(ctas, p.n_blocks) != 1) --ctas;
  }
  // algorithmic HBM bytes of the launch: both operands once, every output (and the aux input) once
  const double io_bytes = 2.0 * (1.0 * o.M + o.N) * o.K + 1.0 * o.M * o.N * p.splits * out.elem_bytes +
                          (out.ptr2 != nullptr ? (Q8 ? 1.0 : 2.0) * o.M * o.N : 0.0) + (out.aux != nullptr ? (Q8 ? 1.0 : 2.0) * o.M * o.N : 0.0);
  const bool prof = b200_prof_gemm_begin(stream, 2.0 * o.M * o.N * o.K, io_bytes);
  if (CG2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * ctas);
    cfg.blockDim = dim3(64 + EW * 32);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaLaunchKernelEx(&cfg, gemm_tn_kernel<Epi, OUT_BYTES, DUAL, AUX, Q8, EW, CG2>, ta, tb, to, to2, tx, p, ep);
  } else {
    launch_pdl(gemm_tn_kernel<Epi, OUT_BYTES, DUAL, AUX, Q8, EW, CG2>, dim3(ctas), dim3(64 + EW * 32), kSmemBytes, stream, ta, tb, to, to2, tx, p, ep);
  }
  if (prof) b200_prof_gemm_end(stream);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// splits actually used for a requested split count (callers size the partial buffer with this)
inline int effective_splits(int K, int splits) {
  int kb = (K + kBlockK - 1) / kBlockK;
  if (splits < 1) splits = 1;
  if (splits > kb) splits = kb;
  int per = (kb + splits - 1) / splits;
  return (kb + per - 1) / per;
}

}  // namespace gemm
