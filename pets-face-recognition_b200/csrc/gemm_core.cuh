// tcgen05 GEMM core for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (both operands K-major, 16-bit),
// fp32 accumulation in TMEM, pluggable epilogue.
//
//   warp 0      TMA producer   : cp.async.bulk.tensor (SWIZZLE_128B boxes of 64 K-elements) -> 4-stage smem ring
//   warp 1      MMA issuer     : one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=block_n, K=16)
//   warps 2..5  epilogue       : tcgen05.ld (32 lanes x 32 columns per warp) -> registers -> Epi::apply -> global
//
// Persistent: grid = min(#tiles, #SMs); every role walks the same static tile sequence
// (tile = blockIdx.x + i * gridDim.x; n fastest so the n-tiles of one A row-block run in the same wave
// and share it through L2).  Two TMEM accumulator stages (2 x 256 columns) let the MMA of tile i+1
// overlap the epilogue of tile i.  block_n is a RUN-TIME multiple of 16 (<= 256): it only enters through
// the B tensor map's box, the instruction descriptor and loop bounds, so one instantiation per epilogue
// serves every layer shape of the network.  Optional split-K (each split writes its own fp32 partial)
// serves the weight-gradient GEMMs whose contraction runs over all tokens.
#pragma once

#include "common.cuh"

namespace gemm {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // 64 x 2 B = one 128-B swizzle row
constexpr int kStages = 4;
constexpr int kMaxBlockN = 256;
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
constexpr int kBBytes = kMaxBlockN * kBlockK * 2; // 32 KB (upper bound; block_n rows are filled)
constexpr int kStageBytes = kABytes + kBBytes;    // 48 KB
constexpr int kTmemCols = 512;                    // 2 accumulator stages x 256 fp32 columns
constexpr int kThreads = 192;
constexpr int kSmemBytes = kStages * kStageBytes + 256 /*barriers*/ + 1024 /*alignment slack*/;

struct CoreParams {
  int M, N;                  // output rows (rows of A) / columns (rows of B)
  int block_n;               // multiple of 16, <= 256
  int m_blocks, n_blocks;
  int k_blocks;              // ceil(K / 64) over the whole contraction
  int splits;                // split-K factor (>= 1)
  int k_blocks_per_split;
  uint32_t idesc;            // tcgen05 instruction descriptor (formats, majors, M=128, N=block_n)
  int mn_major;              // 1: operands are [K, M] / [K, N] row-major (contraction index = row): weight gradients
  int b_chunks;              // mn_major: 64-column chunks of the B tile = ceil(block_n / 64)
};

// instruction descriptor, kind::f16: [4,6) D fmt (1=f32), [7,10) A fmt, [10,13) B fmt (0=f16, 1=bf16),
// [15] A major, [16] B major (0 = K-major), [17,23) N>>3, [24,29) M>>4
inline uint32_t make_idesc(bool is_bf16, int block_n, bool mn_major = false) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (is_bf16 ? 1u : 0u) << 7;
  d |= (is_bf16 ? 1u : 0u) << 10;
  d |= (mn_major ? 1u : 0u) << 15;
  d |= (mn_major ? 1u : 0u) << 16;
  d |= static_cast<uint32_t>(block_n >> 3) << 17;
  d |= static_cast<uint32_t>(kBlockM >> 4) << 24;
  return d;
}

template <class Epi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const CoreParams p, const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B wants 1024-B alignment
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);   // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int num_tiles = p.m_blocks * p.n_blocks * p.splits;
  // K-major: one [128 x 64] A box + one [block_n x 64] B box per stage.  MN-major: the tile is [64 contraction rows x
  // 128 / block_n columns], loaded as [64 x 64] boxes (one 128-B swizzle row = 64 columns) 8 KB apart.
  constexpr uint32_t kChunk = 64 * 64 * 2;
  const uint32_t tx_bytes = p.mn_major ? static_cast<uint32_t>((2 + p.b_chunks) * kChunk)
                                       : static_cast<uint32_t>((kBlockM + p.block_n) * kBlockK * 2);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.n_blocks;
        const int m_blk = (tile / p.n_blocks) % p.m_blocks;
        const int split = tile / (p.n_blocks * p.m_blocks);
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.k_blocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
          const uint32_t sa = smem_base + stage * kStageBytes;
          if (!p.mn_major) {
            tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kBlockK, m_blk * kBlockM);
            tma_load_2d(sa + kABytes, &tmap_b, full_bar(stage), kb * kBlockK, n_blk * p.block_n);
          } else {
            for (int c = 0; c < 2; ++c) tma_load_2d(sa + c * kChunk, &tmap_a, full_bar(stage), m_blk * kBlockM + c * 64, kb * kBlockK);
            for (int c = 0; c < p.b_chunks; ++c)
              tma_load_2d(sa + kABytes + c * kChunk, &tmap_b, full_bar(stage), n_blk * p.block_n + c * 64, kb * kBlockK);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile / (p.n_blocks * p.m_blocks);
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.k_blocks_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kMaxBlockN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          // K-major : SBO = 1024 B between 8-row groups; a K=16 slice is +32 B inside the 128-B swizzle row (+2 in addr>>4)
          // MN-major: LBO = 8 KB between 64-column chunks, SBO = 1024 B between 8-row (contraction) groups; a K=16 slice
          //           is 16 rows = +2048 B (+128 in addr>>4)
          const uint64_t da = p.mn_major ? make_sw128_desc(sa, kChunk, 1024) : make_sw128_desc(sa, 16, 1024);
          const uint64_t db = p.mn_major ? make_sw128_desc(sa + kABytes, kChunk, 1024) : make_sw128_desc(sa + kABytes, 16, 1024);
          const uint32_t kstep = p.mn_major ? 128u : 2u;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) umma_f16(d_tmem, da + kstep * k, db + kstep * k, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar(stage));                 // smem slot reusable once these MMAs retire
          if (kb == kb1 - 1) umma_commit(tfull_bar(acc)); // accumulator complete
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (kb1 <= kb0) umma_commit(tfull_bar(acc));      // degenerate split: nothing to add
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_blk = tile % p.n_blocks;
      const int m_blk = (tile / p.n_blocks) % p.m_blocks;
      const int split = tile / (p.n_blocks * p.m_blocks);
      const int kb0 = split * p.k_blocks_per_split;
      const bool has_k = min(p.k_blocks, kb0 + p.k_blocks_per_split) > kb0;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * kBlockM + q * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * kMaxBlockN);
      Epi::begin_tile(ep, p, row, n_blk * p.block_n, split);
      for (int c = 0; c < p.block_n; c += 32) {
        uint32_t r0[16], r1[16];
        const bool two = (c + 16) < p.block_n;
        tmem_ld16(taddr + c, r0);
        if (two) tmem_ld16(taddr + c + 16, r1);
        tmem_ld_wait();
        if (!has_k) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { r0[i] = 0u; r1[i] = 0u; }
        }
        Epi::apply(ep, p, row, n_blk * p.block_n + c, split, reinterpret_cast<const float(&)[16]>(r0));
        if (two) Epi::apply(ep, p, row, n_blk * p.block_n + c + 16, split, reinterpret_cast<const float(&)[16]>(r1));
      }
      Epi::end_tile(ep, p, row, n_blk * p.block_n, split);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int encode_tmap_2d(CUtensorMap* map, bool is_bf16, const void* ptr, uint64_t inner, uint64_t outer,
                   uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer);

int pick_block_n(int N);

struct Operands {
  const void* a; int lda;     // [M, K] row-major, 16-bit           (mn_major: [K, M] row-major)
  const void* b; int ldb;     // [N, K] row-major, 16-bit           (mn_major: [K, N] row-major)
  int M, N, K;
  bool is_bf16;
  int block_n;                // 0 = auto
  int splits;                 // <= 1 = no split-K
  int max_ctas;               // 0 = #SMs
  bool mn_major = false;      // D = A^T B with the contraction over the ROWS of both operands (weight gradients)
};

template <class Epi>
int launch(const Operands& o, const typename Epi::Params& ep, cudaStream_t stream) {
  B200_REQUIRE(o.M > 0 && o.N > 0 && o.K > 0, "gemm: empty problem M=%d N=%d K=%d", o.M, o.N, o.K);
  B200_REQUIRE(o.lda % 8 == 0 && o.ldb % 8 == 0, "gemm: row pitch must be a multiple of 8 elements (16 B)");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(o.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(o.b) & 15) == 0,
               "gemm: operand base must be 16-B aligned");
  CoreParams p;
  p.M = o.M;
  p.N = o.N;
  p.block_n = o.block_n > 0 ? o.block_n : pick_block_n(o.N);
  B200_REQUIRE(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= kMaxBlockN, "gemm: bad block_n %d", p.block_n);
  p.m_blocks = (o.M + kBlockM - 1) / kBlockM;
  p.n_blocks = (o.N + p.block_n - 1) / p.block_n;
  p.k_blocks = (o.K + kBlockK - 1) / kBlockK;
  p.splits = o.splits > 1 ? o.splits : 1;
  if (p.splits > p.k_blocks) p.splits = p.k_blocks;
  p.k_blocks_per_split = (p.k_blocks + p.splits - 1) / p.splits;
  p.splits = (p.k_blocks + p.k_blocks_per_split - 1) / p.k_blocks_per_split;   // no empty splits
  p.idesc = make_idesc(o.is_bf16, p.block_n, o.mn_major);
  p.mn_major = o.mn_major ? 1 : 0;
  p.b_chunks = (p.block_n + 63) / 64;

  CUtensorMap ta, tb;
  int rc;
  if (!o.mn_major) {
    rc = encode_tmap_2d(&ta, o.is_bf16, o.a, o.K, o.M, o.lda, kBlockK, kBlockM);
    if (rc) return rc;
    rc = encode_tmap_2d(&tb, o.is_bf16, o.b, o.K, o.N, o.ldb, kBlockK, p.block_n);
  } else {
    rc = encode_tmap_2d(&ta, o.is_bf16, o.a, o.M, o.K, o.lda, 64, kBlockK);
    if (rc) return rc;
    rc = encode_tmap_2d(&tb, o.is_bf16, o.b, o.N, o.K, o.ldb, 64, kBlockK);
  }
  if (rc) return rc;

  static bool attr_done = false;   // per instantiation
  if (!attr_done) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done = true;
  }
  const long long tiles = 1LL * p.m_blocks * p.n_blocks * p.splits;
  int ctas = o.max_ctas > 0 ? o.max_ctas : b200_num_sms();
  if (tiles < ctas) ctas = static_cast<int>(tiles);
  const bool prof = b200_prof_gemm_begin(stream, 2.0 * o.M * o.N * o.K);
  gemm_tn_kernel<Epi><<<ctas, kThreads, kSmemBytes, stream>>>(ta, tb, p, ep);
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
