// Gallery matching: cosine(query, gallery) + top-k, the arithmetic behind Recall@K / candR@k
// (reference engine/controller.py:77-91 with similarity_f of configs/dog_fe/fe_dogs_config.py:89-93; the
// query != gallery form is generate_tsv_to_reproduce2.py:63-119).  The reference scores one query at a time in a
// Python loop and fully sorts N-1 scores; here:
//
//  phase 1  cosine_filter_kernel: fp16 unit rows, Q block (128 queries x dim) resident in smem, gallery tiles of
//           256 rows streamed by TMA, tcgen05.mma into two 256-column TMEM accumulators; the epilogue never
//           writes the N x M score matrix: each thread owns one query row, compares its 256 scores per tile
//           against a running threshold (max-tree over 32 scores, one compare) and appends the rare survivors to a
//           per-row candidate list, pruned warp-cooperatively by an exact radix select if it fills.
//           Large galleries run it twice: a first pass over 1/16 of the rows yields, per query, the score of its
//           KP-th best row there - a threshold that can only be BELOW the KP-th best of the whole gallery, hence safe -
//           and the second pass over the remaining rows starts from that threshold, so it appends ~16 KP rows per query
//           instead of flooding and re-pruning its lists in every chunk.
//  phase 2  rerank_kernel: the <= KP candidates per (query, gallery chunk) are re-scored EXACTLY - fp64 cosine from
//           the fp32 embeddings - and sorted by (score desc, gallery index asc): the deterministic order that
//           oracle/rank_oracle.py:topk_spec defines, so indices are bit-exact regardless of fp16 error as long as
//           the true top-k lie in the approximate top-KP (KP = k + 28 slack, fp16 cosine error ~3e-5).
#include <climits>
#include <cstdlib>

#include "common.cuh"

#include "b200_fe.h"
#include "gemm_core.cuh"

namespace {

constexpr int kBM = 128;            // queries per block
constexpr int kBN = 256;            // gallery rows per tile
constexpr int kBK = 64;
constexpr int kMaxKB = 8;           // dim <= 512
constexpr int kBStages = 3;
constexpr int kQSlab = kBM * kBK * 2;        // 16 KB
constexpr int kBStage = kBN * kBK * 2;       // 32 KB
constexpr int kCap = 512;           // candidate list capacity per query row (entries of 8 B)
constexpr int kKP = 128;            // candidates kept per (query, chunk)
constexpr int kThreads = 192;
constexpr int kSmem = kMaxKB * kQSlab + kBStages * kBStage + 256 + 1024;

struct FilterParams {
  long long nq;
  long long g_begin, ng;      // gallery rows [g_begin, ng) are scanned
  const float* tau_init;      // [nq] starting threshold per query, or null (-inf)
  float* tau_out;             // [nq] receives the score of the KP-th best row found (chunks == 1 only), or null
  int list_base, lists;       // this launch fills candidate lists [list_base, list_base + chunks) of `lists` per query
  int kb;                     // dim / 64
  int chunks;                 // gallery chunks
  long long chunk_rows;       // multiple of 256
  int q_blocks;
  long long self_offset;      // gallery row (self_offset + q) is excluded for query q
  int exclude_self;
  uint32_t idesc;
  uint2* scratch;             // [gridDim.x][128][kCap] (score bits, idx)
  int* cand_idx;              // [q_blocks*128][lists][kKP]
  float* cand_score;          // [q_blocks*128][lists][kKP] fp16-GEMM scores of the survivors (approximate)
  int* cand_cnt;              // [q_blocks*128][lists]
};

__device__ __forceinline__ uint32_t fkey(float f) {   // order-preserving float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// Warp-cooperative exact prune of one row's list to its best kKP entries (stable: ties keep list order, which is
// ascending gallery index).  Returns the new count; *tau_out = score of the kKP-th best (or -inf if fewer).
__device__ int prune_list(uint2* list, int n, float* tau_out, int lane) {
  uint32_t key[kCap / 32];
  uint32_t idx[kCap / 32];
#pragma unroll
  for (int i = 0; i < kCap / 32; ++i) {
    const int e = i * 32 + lane;
    if (e < n) { const uint2 v = list[e]; key[i] = fkey(__uint_as_float(v.x)); idx[i] = v.y; }
    else { key[i] = 0u; idx[i] = 0u; }     // key 0 is below every real score's key
  }
  if (n <= kKP) {
    *tau_out = -INFINITY;
    return n;
  }
  uint32_t thr = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = thr | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < kCap / 32; ++i) c += (key[i] >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= kKP) thr = cand;
  }
  int n_gt = 0;
#pragma unroll
  for (int i = 0; i < kCap / 32; ++i) n_gt += (key[i] > thr) ? 1 : 0;
  n_gt = __reduce_add_sync(0xffffffffu, n_gt);
  int eq_left = kKP - n_gt;
  __syncwarp();
  int base = 0;
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kCap / 32; ++i) {
    const bool gt = key[i] > thr;
    const bool eq = key[i] == thr;
    const uint32_t m_eq = __ballot_sync(0xffffffffu, eq);
    const bool eq_keep = eq && (__popc(m_eq & lt_mask) < eq_left);
    const bool keep = gt || eq_keep;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (keep) list[base + __popc(m & lt_mask)] = make_uint2(__float_as_uint(fkey_inv(key[i])), idx[i]);
    base += __popc(m);
    eq_left -= min(eq_left, __popc(m_eq));
  }
  __syncwarp();
  *tau_out = fkey_inv(thr);
  return base;
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose box lands at the same smem offset - and signals the mbarrier at the same offset - in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t smem_dst, const void* tmap, uint32_t bar, int c_inner, int c_outer, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c_inner), "r"(c_outer), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

// CLUSTER CTAs (a thread-block cluster) work on CLUSTER different query blocks against the SAME gallery chunk in lock step:
// each CTA fetches 1/CLUSTER of every gallery tile and TMA-multicasts it into the smem of all of them, so the L2 -> SM
// traffic of the gallery stream (the bound of this kernel: 256 KB per 128 x 256 x 512 tile) drops by CLUSTER.
template <int CLUSTER>
__global__ void __launch_bounds__(kThreads, 1)
cosine_filter_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g, const FilterParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_base = smem_base;
  const uint32_t b_base = smem_base + p.kb * kQSlab;          // Q takes kb slabs; B stages follow
  const uint32_t bar_base = smem_base + kMaxKB * kQSlab + kBStages * kBStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kBStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kBStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kBStages + 2 + a); };
  const uint32_t qfull_bar = bar_base + 8u * (2 * kBStages + 4);
  const uint32_t qempty_bar = bar_base + 8u * (2 * kBStages + 5);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kBStages + 6);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_g);
    for (int s = 0; s < kBStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CLUSTER); }   // every CTA of the cluster releases a slot
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    mbar_init(qfull_bar, 1);
    mbar_init(qempty_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int crank = CLUSTER > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  if (CLUSTER > 1) cluster_sync_all();      // peers' barriers are initialised before anything is multicast into this CTA
  pdl_grid_sync();

  // work units: (group of CLUSTER consecutive query blocks, gallery chunk); CTA `crank` of a cluster takes query block
  // group * CLUSTER + crank (a block past the end scores zero-filled queries and writes nothing)
  const int q_groups = (p.q_blocks + CLUSTER - 1) / CLUSTER;
  const int units = q_groups * p.chunks;
  const int cluster_id = static_cast<int>(blockIdx.x) / CLUSTER, n_clusters = static_cast<int>(gridDim.x) / CLUSTER;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CLUSTER) - 1u);
  constexpr int kSliceRows = kBN / CLUSTER;
  auto tiles_of = [&](int chunk) -> int {
    const long long g0 = p.g_begin + 1LL * chunk * p.chunk_rows;
    const long long g1 = min(p.ng, g0 + p.chunk_rows);
    return static_cast<int>((g1 - g0 + kBN - 1) / kBN);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, qphase = 0;
      for (int unit = cluster_id; unit < units; unit += n_clusters) {
        const int qb = (unit % q_groups) * CLUSTER + crank, chunk = unit / q_groups;
        mbar_wait(qempty_bar, qphase ^ 1u);
        mbar_arrive_expect_tx(qfull_bar, static_cast<uint32_t>(p.kb * kQSlab));
        for (int kb = 0; kb < p.kb; ++kb) tma_load_2d(q_base + kb * kQSlab, &tmap_q, qfull_bar, kb * kBK, qb * kBM);
        qphase ^= 1u;
        const int nt = tiles_of(chunk);
        const long long g0 = p.g_begin + 1LL * chunk * p.chunk_rows;
        for (int t = 0; t < nt; ++t)
          for (int kb = 0; kb < p.kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_arrive_expect_tx(full_bar(stage), kBStage);       // the whole tile: CLUSTER slices, one from each CTA
            if (CLUSTER == 1)
              tma_load_2d(b_base + stage * kBStage, &tmap_g, full_bar(stage), kb * kBK, static_cast<int>(g0 + 1LL * t * kBN));
            else
              tma_load_2d_mc(b_base + stage * kBStage + crank * (kSliceRows * kBK * 2), &tmap_g, full_bar(stage), kb * kBK,
                             static_cast<int>(g0 + 1LL * t * kBN) + crank * kSliceRows, kMask);
            if (++stage == kBStages) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, qphase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int unit = cluster_id; unit < units; unit += n_clusters) {
        const int chunk = unit / q_groups;
        mbar_wait(qfull_bar, qphase);
        qphase ^= 1u;
        tc_fence_after();
        const int nt = tiles_of(chunk);
        for (int t = 0; t < nt; ++t) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kBN);
          for (int kb = 0; kb < p.kb; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint64_t da = make_sw128_desc(q_base + kb * kQSlab, 16, 1024);
            const uint64_t db = make_sw128_desc(b_base + stage * kBStage, 16, 1024);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) umma_f16(d_tmem, da + 2u * k, db + 2u * k, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (CLUSTER == 1) umma_commit(empty_bar(stage));
            else umma_commit_mc(empty_bar(stage), kMask);       // the slot is refilled by all CTAs: tell every producer
            if (kb == p.kb - 1) umma_commit(tfull_bar(acc));
            if (++stage == kBStages) { stage = 0; phase ^= 1u; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
        umma_commit(qempty_bar);     // every MMA that read this unit's Q slabs has retired
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint2* list = p.scratch + (1LL * blockIdx.x * kBM + row) * kCap;
    int acc = 0; uint32_t acc_phase = 0;
    for (int unit = cluster_id; unit < units; unit += n_clusters) {
      const int qb = (unit % q_groups) * CLUSTER + crank, chunk = unit / q_groups;
      const long long qrow = 1LL * qb * kBM + row;
      const bool live = qrow < p.nq;
      const long long self_col = p.exclude_self ? (p.self_offset + qrow) : LLONG_MIN;
      const long long g0 = p.g_begin + 1LL * chunk * p.chunk_rows;
      const long long g1 = min(p.ng, g0 + p.chunk_rows);
      const int nt = tiles_of(chunk);
      float tau = (live && p.tau_init != nullptr) ? p.tau_init[qrow] : -INFINITY;
      int cnt = 0;
      for (int t = 0; t < nt; ++t) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * kBN);
        const long long col0 = g0 + 1LL * t * kBN;
#pragma unroll 1
        for (int c = 0; c < kBN; c += 32) {
          uint32_t r0[16], r1[16];
          tmem_ld16(taddr + c, r0);
          tmem_ld16(taddr + c + 16, r1);
          tmem_ld_wait();
          // one max-tree + one compare per 32 scores; the element-wise scan runs only for the rare groups with a survivor
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
          if (live && mx > tau) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = __uint_as_float(i < 16 ? r0[i] : r1[i - 16]);
              if (v > tau) {
                const long long col = col0 + c + i;
                if (col < g1 && col != self_col && cnt < kCap) { list[cnt] = make_uint2(__float_as_uint(v), static_cast<uint32_t>(col)); ++cnt; }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
        // keep room for a full tile of survivors: prune rows whose list passed half capacity
        uint32_t need = __ballot_sync(0xffffffffu, cnt > kCap - kBN);
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int n = __shfl_sync(0xffffffffu, cnt, src);
          uint2* l = p.scratch + (1LL * blockIdx.x * kBM + q * 32 + src) * kCap;
          float new_tau;
          const int m = prune_list(l, n, &new_tau, lane);
          if (lane == src) { cnt = m; tau = fmaxf(tau, new_tau); }
        }
      }
      // unit done: final prune, then hand the survivors (gallery indices) to phase 2
      {
        uint32_t need = __ballot_sync(0xffffffffu, cnt > kKP);
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int n = __shfl_sync(0xffffffffu, cnt, src);
          uint2* l = p.scratch + (1LL * blockIdx.x * kBM + q * 32 + src) * kCap;
          float new_tau;
          const int m = prune_list(l, n, &new_tau, lane);
          if (lane == src) cnt = m;
        }
        __syncwarp();
        if (qb < p.q_blocks) {
        const long long li = qrow * p.lists + p.list_base + chunk;
        int* dst = p.cand_idx + li * kKP;
        float* dsc = p.cand_score + li * kKP;
        float mn = INFINITY;
        for (int i = 0; i < cnt; ++i) {
          const uint2 e = list[i];
          dst[i] = static_cast<int>(e.y);
          dsc[i] = __uint_as_float(e.x);
          mn = fminf(mn, __uint_as_float(e.x));
        }
        p.cand_cnt[li] = live ? cnt : 0;
        if (p.tau_out != nullptr && live)      // KP-th best of the rows seen so far: a lower bound of the KP-th best overall
          p.tau_out[qrow] = (cnt >= kKP) ? mn : (p.tau_init != nullptr ? p.tau_init[qrow] : -INFINITY);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();      // no CTA may leave while peers can still multicast into it / arrive on its barriers
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------
// phase 2: exact fp64 re-rank + sort.  One CTA per query.
// ---------------------------------------------------------------------------------------------
struct Cand { double score; int idx; int pad; };

__device__ __forceinline__ bool cand_before(const Cand& a, const Cand& b) {   // a ranks ahead of b
  return a.score > b.score || (a.score == b.score && a.idx < b.idx);
}

__device__ void bitonic_sort(Cand* c, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;     // "up" block: best first
          const Cand a = c[i], b = c[l];
          if (up ? cand_before(b, a) : cand_before(a, b)) { c[i] = b; c[l] = a; }
        }
      }
      __syncthreads();
    }
}

__global__ void __launch_bounds__(256) rerank_kernel(const float* __restrict__ q, const double* __restrict__ q_norm,
                                                     const float* __restrict__ g, const double* __restrict__ g_norm, int dim,
                                                     const int* __restrict__ cand_idx, const float* __restrict__ cand_score,
                                                     const int* __restrict__ cand_cnt, int lists, int cap, int k,
                                                     long long g_index_base, int* __restrict__ out_idx, double* __restrict__ out_score) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t sm[];
  Cand* sel = reinterpret_cast<Cand*>(sm);                                                     // [kKP] exact stage
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(sm + kKP * sizeof(Cand));   // [cap] approximate stage
  __shared__ int s_off[64];
  __shared__ int s_total;
  __shared__ int s_cnt[8];
  __shared__ int s_slot;
  const long long qi = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    int t = 0;
    for (int c = 0; c < lists; ++c) { s_off[c] = t; t += cand_cnt[qi * lists + c]; }
    s_total = t;
    s_slot = 0;
  }
  for (int i = threadIdx.x; i < kKP; i += blockDim.x) { sel[i].score = -INFINITY; sel[i].idx = INT_MAX; }
  __syncthreads();
  int total = s_total;
  // unique order-preserving key: (approximate score desc, gallery index asc)  ==  larger key first
  for (int c = 0; c < lists; ++c) {
    const int n = cand_cnt[qi * lists + c];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const long long src = (qi * lists + c) * kKP + i;
      keys[s_off[c] + i] = (static_cast<unsigned long long>(fkey(cand_score[src])) << 32) |
                           static_cast<unsigned long long>(0xffffffffu - static_cast<uint32_t>(cand_idx[src]));
    }
  }
  __syncthreads();
  // level 1: of the survivors of all passes / chunks only the best kKP by approximate (fp16 tensor-core) score can contain
  // the exact top-k (same k + 28 slack argument as inside a chunk): exact 64-bit radix select, O(n) per bit
  unsigned long long thr = 0ULL;
  if (total > kKP) {
#pragma unroll 1
    for (int bit = 63; bit >= 0; --bit) {
      const unsigned long long cand = thr | (1ULL << bit);
      int c = 0;
      for (int i = threadIdx.x; i < total; i += blockDim.x) c += (keys[i] >= cand) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (lane == 0) s_cnt[warp] = c;
      __syncthreads();
      int t = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_cnt[w];
      if (t >= kKP) thr = cand;
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < total; i += blockDim.x)       // keys are unique: exactly min(total, kKP) pass
    if (keys[i] >= thr) sel[atomicAdd(&s_slot, 1)].idx = static_cast<int>(0xffffffffu - static_cast<uint32_t>(keys[i] & 0xffffffffULL));
  __syncthreads();
  total = min(total, kKP);
  // level 2: exact fp64 cosine of those <= kKP candidates from the fp32 embeddings
  const float* qr = q + qi * dim;
  const double nq = fmax(q_norm[qi], 1e-8);
  for (int i = warp; i < total; i += blockDim.x >> 5) {
    const int gi = sel[i].idx;
    const float* gr = g + 1LL * gi * dim;
    double acc = 0.0;
    for (int d = lane * 4; d < dim; d += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(qr + d));
      const float4 b = __ldg(reinterpret_cast<const float4*>(gr + d));
      acc = fma(static_cast<double>(a.x), static_cast<double>(b.x), acc);
      acc = fma(static_cast<double>(a.y), static_cast<double>(b.y), acc);
      acc = fma(static_cast<double>(a.z), static_cast<double>(b.z), acc);
      acc = fma(static_cast<double>(a.w), static_cast<double>(b.w), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sel[i].score = acc / (nq * fmax(g_norm[gi], 1e-8));
  }
  __syncthreads();
  bitonic_sort(sel, kKP);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const bool ok = i < total;
    out_idx[qi * k + i] = ok ? static_cast<int>(sel[i].idx + g_index_base) : -1;
    out_score[qi * k + i] = ok ? sel[i].score : -INFINITY;
  }
}

// merge `lists` pre-scored top lists per query ([lists][nq][k_in]) into one top-k_out
__global__ void __launch_bounds__(256) merge_kernel(const double* __restrict__ scores, const int* __restrict__ idx, long long nq, int lists,
                                                    int k_in, int n_pow2, int k_out, int* __restrict__ out_idx, double* __restrict__ out_score) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t sm[];
  Cand* cands = reinterpret_cast<Cand*>(sm);
  const long long qi = blockIdx.x;
  const int total = lists * k_in;
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
    Cand c; c.score = -INFINITY; c.idx = INT_MAX; c.pad = 0;
    if (i < total) {
      const int l = i / k_in, j = i % k_in;
      const int id = idx[(1LL * l * nq + qi) * k_in + j];
      if (id >= 0) { c.idx = id; c.score = scores[(1LL * l * nq + qi) * k_in + j]; }
    }
    cands[i] = c;
  }
  __syncthreads();
  bitonic_sort(cands, n_pow2);
  for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
    const bool ok = cands[i].idx != INT_MAX;
    out_idx[qi * k_out + i] = ok ? cands[i].idx : -1;
    out_score[qi * k_out + i] = ok ? cands[i].score : -INFINITY;
  }
}

// fp32 rows -> fp16 unit rows + fp64 norms (fixed summation order: lane-strided chains, xor tree)
__global__ void __launch_bounds__(256) gallery_prepare_kernel(const float* __restrict__ x, __half* __restrict__ out, double* __restrict__ norm,
                                                              long long n, int dim) {
  pdl_grid_sync();
  const long long row = (1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + row * dim;
  double ss = 0.0;
  for (int d = lane * 4; d < dim; d += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xr + d));
    ss = fma(static_cast<double>(v.x), static_cast<double>(v.x), ss);
    ss = fma(static_cast<double>(v.y), static_cast<double>(v.y), ss);
    ss = fma(static_cast<double>(v.z), static_cast<double>(v.z), ss);
    ss = fma(static_cast<double>(v.w), static_cast<double>(v.w), ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const double nr = sqrt(ss);
  if (lane == 0) norm[row] = nr;
  const float inv = static_cast<float>(1.0 / fmax(nr, 1e-8));
  for (int d = lane * 4; d < dim; d += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xr + d));
    uint2 o;
    o.x = pack_f16(v.x * inv, v.y * inv);
    o.y = pack_f16(v.z * inv, v.w * inv);
    *reinterpret_cast<uint2*>(out + row * dim + d) = o;
  }
}

__global__ void recall_hits_kernel(const int* __restrict__ top_idx, long long nq, int k_stride, const long long* __restrict__ q_class,
                                   const long long* __restrict__ g_class, const int* __restrict__ ks, int n_ks,
                                   unsigned long long* __restrict__ hits) {
  pdl_grid_sync();
  const long long qi = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  const long long qc = q_class[qi];
  int first = INT_MAX;                       // rank of the first same-class candidate
  for (int j = 0; j < k_stride; ++j) {
    const int id = top_idx[qi * k_stride + j];
    if (id >= 0 && g_class[id] == qc) { first = j; break; }
  }
  for (int i = 0; i < n_ks; ++i)
    if (first < ks[i]) atomicAdd(&hits[i], 1ULL);
}

struct Layout {
  long long scratch, cand_idx, cand_score, cand_cnt, tau, total;
  int q_blocks, lists;
  long long pre0_rows;         // rows [0, pre0_rows): seed pass (floods its lists, tiny)
  long long pre_rows;          // rows [pre0_rows, pre_rows): threshold pass (0 = single pass over everything)
  int chunks;                  // gallery chunks of the main pass
  long long chunk_rows;
  int ctas_pre, ctas;
  long long ng;                // gallery rows
  float* tau_ptr;              // [q_blocks*128] thresholds handed from pass to pass
};

// chunks for `rows` gallery rows: enough (q_block, chunk) units to fill the SMs ~6 times, chunks >= 16 tiles, and the unit
// count close to a multiple of the CTA count (the units are equal-sized: a ragged last wave is pure loss)
void pick_chunks(long long rows, int q_blocks, int sms, int* chunks, long long* chunk_rows) {
  const long long tiles = (rows + kBN - 1) / kBN;
  long long want = (6LL * sms + q_blocks - 1) / q_blocks;
  const long long cap = std::max<long long>(1, std::min<long long>(tiles / 16, 31));
  want = std::max<long long>(1, std::min(want, cap));
  long long best = want;
  double best_waste = 1e9;
  for (long long c = want; c <= std::min(cap, want * 2); ++c) {
    const long long units = c * q_blocks, ctas = std::min<long long>(units, sms);
    const double waste = static_cast<double>((units + ctas - 1) / ctas * ctas - units) / units;
    if (waste < best_waste - 1e-9) { best_waste = waste; best = c; }
  }
  const long long tpc = (tiles + best - 1) / best;
  *chunk_rows = tpc * kBN;
  *chunks = static_cast<int>((tiles + tpc - 1) / tpc);
}

Layout plan_layout(long long nq, long long ng) {
  Layout L;
  L.q_blocks = static_cast<int>((nq + kBM - 1) / kBM);
  const int sms = b200_num_sms();
  L.pre_rows = L.pre0_rows = 0;
  if (ng >= 16LL * 4096) {                                                   // thresholds from the first 1/16 of the rows,
    L.pre_rows = (ng / 16 + kBN - 1) / kBN * kBN;                            // itself seeded from the first 1024 rows
    L.pre0_rows = 4 * kBN;
  }
  pick_chunks(ng - L.pre_rows, L.q_blocks, sms, &L.chunks, &L.chunk_rows);
  L.lists = L.chunks + (L.pre_rows ? 2 : 0);
  L.ctas_pre = std::min(L.q_blocks, sms);
  L.ctas = static_cast<int>(std::min<long long>(1LL * L.q_blocks * L.chunks, sms));
  long long off = 0;
  auto take = [&](long long bytes) { long long o = off; off = (off + bytes + 255) / 256 * 256; return o; };
  L.scratch = take(1LL * sms * kBM * kCap * 8);
  L.cand_idx = take(1LL * L.q_blocks * kBM * L.lists * kKP * 4);
  L.cand_score = take(1LL * L.q_blocks * kBM * L.lists * kKP * 4);
  L.cand_cnt = take(1LL * L.q_blocks * kBM * L.lists * 4);
  L.tau = take(1LL * L.q_blocks * kBM * 4);
  L.total = off;
  return L;
}

constexpr int kCluster = 4;

template <int CLUSTER>
int launch_one(const CUtensorMap& tq, const CUtensorMap& tg, const FilterParams& p, int ctas, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(cosine_filter_kernel<CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  int n = 1;
  if (CLUSTER > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = CLUSTER; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    n = 2;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  B200_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cosine_filter_kernel<CLUSTER>, tq, tg, p));
  b200_count_launch();
  return B200_OK;
}

// how many clusters of kCluster CTAs can be resident at once (the GPC layout strands a few SMs for clusters of 4)
int max_clusters() {
  static int cached = -1;
  if (cached < 0) {
    cudaFuncSetAttribute(cosine_filter_kernel<kCluster>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(b200_num_sms() / kCluster * kCluster);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, cosine_filter_kernel<kCluster>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 0; }
    cached = n;
  }
  return cached;
}

// the three passes of one top-k call; clusters when there are enough query blocks to fill them
int launch_filter(const CUtensorMap& tq, const CUtensorMap& tg, const CUtensorMap& tg_slice, FilterParams p, const Layout& L, cudaStream_t st) {
  const int sms = b200_num_sms();
  const int ncl = max_clusters();
  static const bool env_off = [] { const char* e = getenv("B200_GALLERY_CLUSTER"); return e != nullptr && e[0] == '0'; }();
  const bool use_cluster = !env_off && ncl > 0 && L.q_blocks >= 2 * kCluster;
  auto run = [&](int units_single) -> int {
    if (use_cluster) {
      const int q_groups = (L.q_blocks + kCluster - 1) / kCluster;
      const int units = q_groups * p.chunks;
      return launch_one<kCluster>(tq, tg_slice, p, std::min(units, ncl) * kCluster, st);
    }
    return launch_one<1>(tq, tg, p, std::min(units_single, sms), st);
  };
  int rc;
  if (L.pre_rows > 0) {      // threshold passes: one chunk each, exact streaming top-KP of the rows they scan
    p.g_begin = 0; p.ng = L.pre0_rows; p.chunks = 1; p.chunk_rows = L.pre0_rows; p.list_base = 0;
    p.tau_init = nullptr; p.tau_out = L.tau_ptr;
    if ((rc = run(L.q_blocks))) return rc;
    p.g_begin = L.pre0_rows; p.ng = L.pre_rows; p.chunk_rows = L.pre_rows - L.pre0_rows; p.list_base = 1;
    p.tau_init = L.tau_ptr; p.tau_out = L.tau_ptr;
    if ((rc = run(L.q_blocks))) return rc;
  }
  p.g_begin = L.pre_rows; p.ng = L.ng; p.chunks = L.chunks; p.chunk_rows = L.chunk_rows; p.list_base = L.pre_rows ? 2 : 0;
  p.tau_init = L.pre_rows ? L.tau_ptr : nullptr; p.tau_out = nullptr;
  return run(L.q_blocks * L.chunks);
}

}  // namespace

extern "C" int b200_gallery_prepare(const float* emb, void* unit_f16, double* norm, long long n, int dim, void* stream) {
  B200_REQUIRE(dim % 4 == 0, "gallery_prepare: dim must be a multiple of 4");
  if (n == 0) return B200_OK;
  const long long blocks = (n * 32 + 255) / 256;
  launch_pdl(gallery_prepare_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      emb, reinterpret_cast<__half*>(unit_f16), norm, n, dim);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" long long b200_cosine_topk_workspace_bytes(long long nq, long long ng, int dim, int k) {
  (void)dim; (void)k;
  if (nq <= 0 || ng <= 0) return 256;
  return plan_layout(nq, ng).total;
}

extern "C" int b200_cosine_topk(const float* q, const void* q_unit_f16, const double* q_norm, long long nq, const float* g,
                                const void* g_unit_f16, const double* g_norm, long long ng, int dim, int k,
                                long long exclude_self_offset, long long g_index_base, int* out_idx, double* out_score,
                                void* workspace, long long workspace_bytes, void* stream) {
  B200_REQUIRE(dim % 64 == 0 && dim >= 64 && dim <= 64 * kMaxKB, "cosine_topk: dim=%d must be a multiple of 64, <= 512", dim);
  B200_REQUIRE(k >= 1 && k <= kKP - 28, "cosine_topk: k=%d must be in [1, %d] (KP=%d candidates with 28 slack)", k, kKP - 28, kKP);
  B200_REQUIRE(ng < (1LL << 31) && nq < (1LL << 31), "cosine_topk: more than 2^31 rows");
  if (nq == 0) return B200_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (ng == 0) {
    B200_CHECK_CUDA(cudaMemsetAsync(out_idx, 0xff, sizeof(int) * nq * k, st));
    return b200_set_error(B200_ERR_INVALID, "cosine_topk: empty gallery");
  }
  Layout L = plan_layout(nq, ng);
  B200_REQUIRE(L.lists <= 64, "cosine_topk: too many candidate lists");
  if (workspace_bytes < L.total)
    return b200_set_error(B200_ERR_WORKSPACE, "cosine_topk: workspace %lld < required %lld bytes", workspace_bytes, L.total);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  FilterParams p;
  p.nq = nq; p.kb = dim / kBK; p.q_blocks = L.q_blocks; p.lists = L.lists;
  p.exclude_self = exclude_self_offset != B200_NO_EXCLUDE;
  p.self_offset = p.exclude_self ? exclude_self_offset : 0;
  p.idesc = gemm::make_idesc(false, kBN);
  p.scratch = reinterpret_cast<uint2*>(ws + L.scratch);
  p.cand_idx = reinterpret_cast<int*>(ws + L.cand_idx);
  p.cand_score = reinterpret_cast<float*>(ws + L.cand_score);
  p.cand_cnt = reinterpret_cast<int*>(ws + L.cand_cnt);
  L.ng = ng;
  L.tau_ptr = reinterpret_cast<float*>(ws + L.tau);
  CUtensorMap tq, tg;
  int rc = gemm::encode_tmap_2d(&tq, false, q_unit_f16, dim, nq, dim, kBK, kBM);
  if (rc) return rc;
  rc = gemm::encode_tmap_2d(&tg, false, g_unit_f16, dim, ng, dim, kBK, kBN);
  if (rc) return rc;
  CUtensorMap tg_slice;      // one cluster CTA's share of a gallery tile
  rc = gemm::encode_tmap_2d(&tg_slice, false, g_unit_f16, dim, ng, dim, kBK, kBN / kCluster);
  if (rc) return rc;
  rc = launch_filter(tq, tg, tg_slice, p, L, st);
  if (rc) return rc;
  const int n_pow2 = L.lists * kKP;        // candidate capacity per query
  const int smem2 = kKP * static_cast<int>(sizeof(Cand)) + n_pow2 * 8;
  if (smem2 > 48 * 1024) B200_CHECK_CUDA(cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
  launch_pdl(rerank_kernel, dim3(static_cast<unsigned>(nq)), dim3(256), smem2, st, q, q_norm, g, g_norm, dim, p.cand_idx, p.cand_score, p.cand_cnt, L.lists, n_pow2, k,
                                                               g_index_base, out_idx, out_score);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_topk_merge(const double* scores, const int* idx, long long nq, int lists, int k_in, int k_out, int* out_idx,
                               double* out_score, void* stream) {
  B200_REQUIRE(lists >= 1 && k_in >= 1 && k_out >= 1 && k_out <= lists * k_in, "topk_merge: bad sizes");
  if (nq == 0) return B200_OK;
  int n_pow2 = 1;
  while (n_pow2 < lists * k_in) n_pow2 <<= 1;
  const int smem = n_pow2 * static_cast<int>(sizeof(Cand));
  B200_REQUIRE(smem <= 200 * 1024, "topk_merge: lists * k_in too large");
  if (smem > 48 * 1024) B200_CHECK_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  launch_pdl(merge_kernel, dim3(static_cast<unsigned>(nq)), dim3(256), smem, reinterpret_cast<cudaStream_t>(stream), scores, idx, nq, lists, k_in, n_pow2, k_out,
                                                                                              out_idx, out_score);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_recall_hits(const int* top_idx, long long nq, int k_stride, const long long* q_class, const long long* g_class,
                                const int* ks, int n_ks, unsigned long long* hits, void* stream) {
  if (nq == 0) return B200_OK;
  launch_pdl(recall_hits_kernel, dim3(static_cast<unsigned>((nq + 255) / 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), top_idx, nq, k_stride, q_class,
                                                                                                              g_class, ks, n_ks, hits);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
