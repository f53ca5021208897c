// Gallery matching: cosine(query, gallery) + top-k, the arithmetic behind Recall@K / candR@k
// (reference engine/controller.py:77-91 with similarity_f of configs/dog_fe/fe_dogs_config.py:89-93; the
// query != gallery form is generate_tsv_to_reproduce2.py:63-119).  The reference scores one query at a time in a
// Python loop and fully sorts N-1 scores; here:
//
//  phase 1  cosine_filter_kernel: fp16 unit rows, Q block (128 queries x dim) resident in smem, gallery tiles of
//           256 rows streamed by TMA, tcgen05.mma into two 256-column TMEM accumulators; the epilogue never
//           writes the N x M score matrix: eight epilogue warps (a TMEM lane quadrant x a 128-column part each) let every
//           thread own one query row: it compares its scores against a running threshold (max-tree over 32 scores, one
//           compare), turns the rare survivors into a bit mask and appends them to a per-(row, half) candidate list,
//           pruned warp-cooperatively by a radix select when it fills.
//           Large galleries run it twice.  A first "witness" pass over the first 8192 rows keeps no lists at all: per
//           query it takes the maximum of each group of 64 scores and the minimum of those 128 maxima - 128 distinct rows
//           score at least that much, so it can only be BELOW the KP-th best of the whole gallery, hence a safe starting
//           threshold (it passes ~9 % of the rows).  The second pass over all rows starts from it instead of flooding.
//  phase 2  rerank_kernel: the <= KP candidates per (query, gallery chunk) are re-scored EXACTLY - fp64 cosine from
//           the fp32 embeddings - and sorted by (score desc, gallery index asc): the deterministic order that
//           oracle/rank_oracle.py:topk_spec defines, so indices are bit-exact regardless of fp16 error as long as
//           the true top-k lie in the approximate top-KP (KP = k + 28 slack, fp16 cosine error ~3e-5).
//  frame    Embeddings of one domain are concentrated (cosines of 0.99+ between unrelated images for an untrained / early
//           backbone): plain fp16 unit rows have no resolution left there.  The tensor-core pass therefore runs in a frame
//           fitted to the gallery: mu = mean unit gallery row, H = the Householder reflection that maps mu / |mu| onto the
//           first axis (orthogonal: dot products are unchanged), and
//               gallery row  G = scale * [ (H g^)_1 - |mu| , (H g^)_2.. ]   ( = H (g^ - mu) )        query  Q = H q^
//           so that Q . G / scale = q^ . (g^ - mu) = cos(q, g) - q^ . mu: the cosine up to a per-query constant - the ranking
//           is untouched - while every large number has left the fp16 operands: all of the gallery vector and all but one
//           coordinate of the query are of the size of the embeddings' SPREAD, and so are their rounding errors.
//  phase 3  CERTIFICATE: every candidate ever dropped had an approximate score <= a_cut (the largest threshold any of the
//           query's lists pruned with / the level-1 cut).  With E = the rigorous bound of |approximate - exact| from the
//           measured fp16 residual norms, a_cut + E < (exact k-th score) proves that no dropped row belongs to the top-k.
//           Queries that cannot be certified (near-duplicate galleries: hundreds of rows closer than fp16 resolution) are
//           re-done by exact_topk_kernel, an exact fp64 scan of the whole gallery - slow, rare, never wrong.
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
constexpr int kBStages = 3;                  // single-CTA mode: three 32-KB gallery stages; CTA-pair mode: six 16-KB half tiles
constexpr int kQSlab = kBM * kBK * 2;        // 16 KB
constexpr int kBStage = kBN * kBK * 2;       // 32 KB
constexpr int kParts = 2;            // column parts of a tile, each with its own epilogue warps and candidate lists (see kEpiWarps)
constexpr int kHalf = kBN / kParts;  // tile columns per epilogue warp
constexpr int kCap = 384;           // candidate list capacity per (query row, column part) (entries of 8 B)
constexpr int kKP = 128;            // candidates kept per (query, chunk, column part)
constexpr int kEpiWarps = 4 * kParts;   // a TMEM lane quadrant x a column part each.  Measured (r02zf): kParts = 4 (sixteen warps, 64-column
                                        // parts, twice the lists) is SLOWER - 11.5 vs 11.1 ms at 50 k x 125 k, 53.8 vs 50.9 ms at 50 k x 1 M:
                                        // every list keeps its own KP candidates, so the survivors (and prunes) nearly double
constexpr int kThreads = 64 + 32 * kEpiWarps;     // TMA warp, MMA warp, epilogue warps
constexpr int kWitnessRows = 32 * kBN;            // rows of the witness pass: kParts parts x (32 * kHalf / 64) groups of 64 = 128 witnesses
constexpr int kSel2 = 512;           // candidates re-scored exactly in the second certification pass (see rerank_kernel)
constexpr int kMaxBStages = 2 * kBStages;
constexpr int kSmem = kMaxKB * kQSlab + kBStages * kBStage + 256 + 1024;

struct FilterParams {
  long long nq;
  long long g_begin, ng;      // gallery rows [g_begin, ng) are scanned
  const float* tau_init;      // [nq][kParts] starting threshold per query = min over the parts, or null (-inf)
  float* tau_out;             // witness pass: [nq][kParts] receives each column part's minimum of group maxima
  uint32_t* tau_shared;       // [nq] order-preserving keys (0 = unset): best threshold any (chunk, half) list of the query has
                              // reached - every list's KP-th best is a lower bound of the query's KP-th best overall, so
                              // the lists of one query, scanned concurrently by different warps / SMs, tighten each other
  int witness;                // 1: witness pass (no candidate lists)
  int lists;                  // candidate lists per query = kParts * chunks; (chunk, part) fills list kParts * chunk + part
  int kb;                     // dim / 64
  int chunks;                 // gallery chunks
  long long chunk_rows;       // multiple of 256
  // ragged last wave: the last `q_groups - full_groups` query groups are cut into tail_chunks chunks of tail_chunk_rows rows each
  // (0: every group has p.chunks chunks) - see plan_layout
  int full_groups, tail_chunks;
  long long tail_chunk_rows;
  int q_blocks;
  long long self_offset;      // gallery row (self_offset + q) is excluded for query q
  int exclude_self;
  uint32_t idesc;
  uint32_t wait_ns;           // suspend-time hint of the producer / MMA-issuer barrier waits
  uint2* scratch;             // [gridDim.x][2][128][kCap] (score bits, idx)
  int* cand_idx;              // [q_blocks*128][lists][kKP]
  float* cand_score;          // [q_blocks*128][lists][kKP] fp16-GEMM scores of the survivors (approximate)
  int* cand_cnt;              // [q_blocks*128][lists]
  float* list_tau;            // [q_blocks*128][lists] final threshold of every list: nothing it dropped scored above it
};

__device__ __forceinline__ uint32_t fkey(float f) {   // order-preserving float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// Warp-cooperative prune of one row's list to its best kKP entries (stable: ties keep list order, which is ascending
// gallery index) by a radix select, two key bits per round.  EXACT = false may stop once the upper 16 key bits are fixed
// and at most kKP + 64 entries share that prefix or beat it: it keeps all of those (the threshold is the prefix's lower
// edge, still a true lower bound of the kKP-th best) and saves the dependent rounds that would only split ties.
// Returns the new count; *tau_out = the threshold below which entries were dropped (-inf if nothing was).
template <bool EXACT>
__device__ int prune_list(uint2* list, int n, float* tau_out, int lane) {
  constexpr int R = kCap / 32;
  uint32_t key[R];
  uint32_t idx[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int e = i * 32 + lane;
    if (e < n) { const uint2 v = list[e]; key[i] = fkey(__uint_as_float(v.x)); idx[i] = v.y; }
    else { key[i] = 0u; idx[i] = 0u; }     // key 0 is below every real score's key
  }
  if (n <= kKP) {
    *tau_out = -INFINITY;
    return n;
  }
  uint32_t thr = 0;
  int cge = n;             // entries with key >= thr
  bool keep_all_ge = false;
#pragma unroll 1
  for (int bit = 30; bit >= 0; bit -= 2) {
    const uint32_t c1 = thr | (1u << bit), c2 = thr | (2u << bit), c3 = thr | (3u << bit);
    int n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      n1 += (key[i] >= c1) ? 1 : 0;
      n2 += (key[i] >= c2) ? 1 : 0;
      n3 += (key[i] >= c3) ? 1 : 0;
    }
    n1 = __reduce_add_sync(0xffffffffu, n1);
    n2 = __reduce_add_sync(0xffffffffu, n2);
    n3 = __reduce_add_sync(0xffffffffu, n3);
    if (n3 >= kKP) { thr = c3; cge = n3; }
    else if (n2 >= kKP) { thr = c2; cge = n2; }
    else if (n1 >= kKP) { thr = c1; cge = n1; }
    if (!EXACT && bit <= 16 && thr != 0u && cge <= kKP + 64) { keep_all_ge = true; break; }
  }
  int eq_left = 0;
  if (!keep_all_ge) {
    int n_gt = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) n_gt += (key[i] > thr) ? 1 : 0;
    n_gt = __reduce_add_sync(0xffffffffu, n_gt);
    eq_left = kKP - n_gt;
  }
  __syncwarp();
  int base = 0;
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const bool gt = keep_all_ge ? (key[i] >= thr) : (key[i] > thr);
    const bool eq = !keep_all_ge && key[i] == thr;
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

// value i of the 32 held in two 16-register arrays, by a 5-level select tree (a dynamic register index would spill)
__device__ __forceinline__ float pick32(const uint32_t (&a)[16], const uint32_t (&b)[16], int i) {
  uint32_t t[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) t[j] = (i & 16) ? b[j] : a[j];
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] = (i & 8) ? t[j + 8] : t[j];
#pragma unroll
  for (int j = 0; j < 4; ++j) t[j] = (i & 4) ? t[j + 4] : t[j];
#pragma unroll
  for (int j = 0; j < 2; ++j) t[j] = (i & 2) ? t[j + 2] : t[j];
  return __uint_as_float((i & 1) ? t[1] : t[0]);
}

// PAIR = 2: two CTAs of a cluster (one TPC) work on two consecutive query blocks against the SAME gallery chunk as one
// M = 256 MMA (cta_group::2, see gemm_core.cuh: CG2): each CTA keeps its own resident query block but only HALF of every
// gallery tile, so the shared-memory traffic per SM - the bound of the single-CTA kernel: (128 + 256) x 64 x 2 B read and
// 256 x 64 x 2 B written per K block, 156 B per clock against the 128 B / clock an SM's shared memory delivers - drops to 94.
template <int PAIR>
__global__ void __launch_bounds__(kThreads, 1)
cosine_filter_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g, const FilterParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int kStages = PAIR * kBStages;                    // same bytes either way
  constexpr int kStageBytes = kBStage / PAIR;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_base = smem_base;
  const uint32_t b_base = smem_base + p.kb * kQSlab;          // Q takes kb slabs; B stages follow
  const uint32_t bar_base = smem_base + kMaxKB * kQSlab + kBStages * kBStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxBStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxBStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxBStages + 2 + a); };
  const uint32_t qfull_bar = bar_base + 8u * (2 * kMaxBStages + 4);
  const uint32_t qempty_bar = bar_base + 8u * (2 * kMaxBStages + 5);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxBStages + 6);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = PAIR > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = crank == 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_g);
    // pair: the leader's full / qfull barriers count both producers (and both CTAs' bytes), its tempty both CTAs' epilogues
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), PAIR); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), PAIR * kEpiWarps); }
    mbar_init(qfull_bar, PAIR);
    mbar_init(qempty_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    if (PAIR > 1) { tmem_alloc_cg2(tmem_slot, 512); tmem_relinquish_cg2(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR > 1) cluster_sync_all();      // the peer's barriers are initialised before anything arrives on them from this CTA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_grid_sync();

  // work units: (group of PAIR consecutive query blocks, gallery chunk); CTA `crank` of a pair takes query block
  // group * PAIR + crank (a block past the end scores zero-filled queries and writes nothing)
  const int q_groups = (p.q_blocks + PAIR - 1) / PAIR;
  const int full_groups = p.tail_chunks > 0 ? p.full_groups : q_groups;
  const int tail_groups = q_groups - full_groups;
  const int units_full = full_groups * p.chunks;
  const int units = units_full + tail_groups * p.tail_chunks;
  const int cluster_id = static_cast<int>(blockIdx.x) / PAIR, n_clusters = static_cast<int>(gridDim.x) / PAIR;
  constexpr int kSliceRows = kBN / PAIR;
  struct Unit { int qg, chunk, nt; long long g0, g1; };
  auto unit_of = [&](int unit) -> Unit {
    Unit u;
    long long rows;
    if (unit < units_full) { u.chunk = unit / full_groups; u.qg = unit - u.chunk * full_groups; rows = p.chunk_rows; }
    else { const int v = unit - units_full; u.chunk = v / tail_groups; u.qg = full_groups + v - u.chunk * tail_groups; rows = p.tail_chunk_rows; }
    u.g0 = p.g_begin + 1LL * u.chunk * rows;
    u.g1 = min(p.ng, u.g0 + rows);
    u.nt = u.g1 > u.g0 ? static_cast<int>((u.g1 - u.g0 + kBN - 1) / kBN) : 0;
    return u;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, qphase = 0;
      for (int unit = cluster_id; unit < units; unit += n_clusters) {
        const Unit u = unit_of(unit);
        const int qb = u.qg * PAIR + crank;
        mbar_wait_backoff(qempty_bar, qphase ^ 1u, p.wait_ns);
        if (PAIR == 1) {
          mbar_arrive_expect_tx(qfull_bar, static_cast<uint32_t>(p.kb * kQSlab));
          for (int kb = 0; kb < p.kb; ++kb) tma_load_2d(q_base + kb * kQSlab, &tmap_q, qfull_bar, kb * kBK, qb * kBM);
        } else {
          if (leader) mbar_arrive_expect_tx(qfull_bar, static_cast<uint32_t>(2 * p.kb * kQSlab));
          for (int kb = 0; kb < p.kb; ++kb) tma_load_2d_cg2(q_base + kb * kQSlab, &tmap_q, qfull_bar, kb * kBK, qb * kBM);
          if (!leader) mbar_arrive_remote(qfull_bar, 0u);
        }
        qphase ^= 1u;
        const int nt = u.nt;
        const long long g0 = u.g0;
        for (int t = 0; t < nt; ++t)
          for (int kb = 0; kb < p.kb; ++kb) {
            mbar_wait_backoff(empty_bar(stage), phase ^ 1u, p.wait_ns);
            if (PAIR == 1) {
              mbar_arrive_expect_tx(full_bar(stage), kBStage);
              tma_load_2d(b_base + stage * kStageBytes, &tmap_g, full_bar(stage), kb * kBK, static_cast<int>(g0 + 1LL * t * kBN));
            } else {
              if (leader) mbar_arrive_expect_tx(full_bar(stage), kBStage);       // the whole tile: one half from each CTA
              tma_load_2d_cg2(b_base + stage * kStageBytes, &tmap_g, full_bar(stage), kb * kBK,
                              static_cast<int>(g0 + 1LL * t * kBN) + crank * kSliceRows);
              if (!leader) mbar_arrive_remote(full_bar(stage), 0u);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      int stage = 0; uint32_t phase = 0, qphase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int unit = cluster_id; unit < units; unit += n_clusters) {
        mbar_wait_backoff(qfull_bar, qphase, p.wait_ns);
        qphase ^= 1u;
        tc_fence_after();
        const int nt = unit_of(unit).nt;
        for (int t = 0; t < nt; ++t) {
          mbar_wait_backoff(tempty_bar(acc), acc_phase ^ 1u, p.wait_ns);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kBN);
          for (int kb = 0; kb < p.kb; ++kb) {
            mbar_wait_backoff(full_bar(stage), phase, p.wait_ns);
            tc_fence_after();
            const uint64_t da = make_sw128_desc(q_base + kb * kQSlab, 16, 1024);
            const uint64_t db = make_sw128_desc(b_base + stage * kStageBytes, 16, 1024);
            if (PAIR == 1) {
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) umma_f16(d_tmem, da + 2u * k, db + 2u * k, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
              umma_commit(empty_bar(stage));
              if (kb == p.kb - 1) umma_commit(tfull_bar(acc));
            } else {
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) umma_f16_cg2(d_tmem, da + 2u * k, db + 2u * k, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
              umma_commit_cg2(empty_bar(stage));                 // the slot is free again in both CTAs
              if (kb == p.kb - 1) umma_commit_cg2(tfull_bar(acc));
            }
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
        if (PAIR == 1) umma_commit(qempty_bar);     // every MMA that read this unit's Q slabs has retired
        else umma_commit_cg2(qempty_bar);
      }
      if (PAIR > 1)        // the peer's last arrivals on this CTA's barriers must have landed before it may exit
        for (int a2 = 0; a2 < 2; ++a2) {
          mbar_wait_backoff(tempty_bar(acc), acc_phase ^ 1u, p.wait_ns);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
    }
  } else {
    const int q = warp & 3;                     // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;           // column part of every tile (the name dates from two parts)
    const int row = q * 32 + lane;
    uint2* warp_lists = p.scratch + ((1LL * blockIdx.x * kParts + half) * kBM + q * 32) * kCap;
    uint2* list = warp_lists + 1LL * lane * kCap;
    int acc = 0; uint32_t acc_phase = 0;
    for (int unit = cluster_id; unit < units; unit += n_clusters) {
      const Unit u = unit_of(unit);
      const int qb = u.qg * PAIR + crank, chunk = u.chunk;
      const long long qrow = 1LL * qb * kBM + row;
      const bool live = qrow < p.nq;
      const long long self_col = p.exclude_self ? (p.self_offset + qrow) : -(1LL << 62);
      const long long g0 = u.g0, g1 = u.g1;
      const int nt = u.nt;
      float tau = -INFINITY;
      if (live && p.tau_init != nullptr) {
        tau = p.tau_init[kParts * qrow];
#pragma unroll
        for (int j = 1; j < kParts; ++j) tau = fminf(tau, p.tau_init[kParts * qrow + j]);
      }
      uint32_t* tau_g = (live && !p.witness) ? p.tau_shared + qrow : nullptr;
      uint32_t shared_key = tau_g != nullptr ? __ldcg(tau_g) : 0u;       // refreshed once per tile, applied one tile later
      int cnt = 0;
      float wmin = INFINITY, grp = -INFINITY;        // witness pass: min over 64-column groups of the group maximum
      for (int t = 0; t < nt; ++t) {
        if (shared_key != 0u) tau = fmaxf(tau, fkey_inv(shared_key));
        if (tau_g != nullptr) shared_key = __ldcg(tau_g);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * kBN + half * kHalf);
        const long long col0 = g0 + 1LL * t * kBN + half * kHalf;
#pragma unroll 1
        for (int c = 0; c < kHalf; c += 32) {
          uint32_t r0[16], r1[16];
          tmem_ld16(taddr + c, r0);
          tmem_ld16(taddr + c + 16, r1);
          tmem_ld_wait();
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
          const long long cbase = col0 + c;
          const long long ds = self_col - cbase;
          if (p.witness) {
            if (static_cast<unsigned long long>(ds) < 32ULL) {      // the query itself is no witness: maximum of the other 31
              mx = -INFINITY;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i != static_cast<int>(ds)) mx = fmaxf(mx, __uint_as_float(i < 16 ? r0[i] : r1[i - 16]));
            }
            grp = fmaxf(grp, mx);
            if (c & 32) { wmin = fminf(wmin, grp); grp = -INFINITY; }
            continue;
          }
          // one max-tree + one compare per 32 scores.  Some lane of the warp nearly always has a survivor, so the slow
          // path is branch-free up to the (usually single) append: a bit mask of the survivors, then one loop over its bits
          if (live && mx > tau) {
            uint32_t mask = 0u;
#pragma unroll
            for (int i = 0; i < 32; ++i) mask |= (__uint_as_float(i < 16 ? r0[i] : r1[i - 16]) > tau) ? (1u << i) : 0u;
            const bool single = (mask & (mask - 1u)) == 0u;       // then the survivor is the maximum itself
            const long long rem = g1 - cbase;
            if (static_cast<unsigned long long>(ds) < 32ULL) mask &= ~(1u << static_cast<int>(ds));
            if (rem < 32) mask &= (rem <= 0) ? 0u : ((1u << static_cast<int>(rem)) - 1u);
            if (__popc(mask) <= 6) {
              while (mask != 0u && cnt < kCap) {
                const int i = __ffs(mask) - 1;
                mask &= mask - 1u;
                const float v = single ? mx : pick32(r0, r1, i);
                list[cnt] = make_uint2(__float_as_uint(v), static_cast<uint32_t>(cbase + i));
                ++cnt;
              }
            } else {       // a flood (unseeded threshold): walk the registers in order instead of selecting each survivor
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (((mask >> i) & 1u) && cnt < kCap) {
                  list[cnt] = make_uint2(i < 16 ? r0[i] : r1[i - 16], static_cast<uint32_t>(cbase + i));
                  ++cnt;
                }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR > 1 && !leader) mbar_arrive_remote(tempty_bar(acc), 0u);     // the leader's MMA thread waits for both CTAs
          else mbar_arrive(tempty_bar(acc));
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
        // keep room for a full half tile of survivors
        uint32_t need = __ballot_sync(0xffffffffu, cnt > kCap - kHalf);
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int n = __shfl_sync(0xffffffffu, cnt, src);
          float new_tau;
          const int m = prune_list<false>(warp_lists + 1LL * src * kCap, n, &new_tau, lane);
          if (lane == src) {
            cnt = m;
            if (new_tau > tau) { tau = new_tau; atomicMax(tau_g, fkey(new_tau)); }
          }
        }
      }
      if (p.witness) {
        // 1e-6 below the weakest witness: the filter compares with >, and a tie with the threshold must survive
        if (live) p.tau_out[kParts * qrow + half] = wmin - 1e-6f;
        continue;
      }
      // unit done: exact final prune, then hand the survivors (gallery indices) to phase 2, one row at a time, coalesced
      __syncwarp();
      uint32_t need = __ballot_sync(0xffffffffu, cnt > kKP);
      while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const int n = __shfl_sync(0xffffffffu, cnt, src);
        float new_tau;
        const int m = prune_list<true>(warp_lists + 1LL * src * kCap, n, &new_tau, lane);
        if (lane == src) { cnt = m; tau = fmaxf(tau, new_tau); }
      }
      __syncwarp();
      if (qb < p.q_blocks) {
        for (int r = 0; r < 32; ++r) {
          const int n = __shfl_sync(0xffffffffu, cnt, r);
          const long long rq = 1LL * qb * kBM + q * 32 + r;
          const long long li = rq * p.lists + kParts * chunk + half;
          const uint2* src = warp_lists + 1LL * r * kCap;
          for (int i = lane; i < n; i += 32) {
            const uint2 e = src[i];
            p.cand_idx[li * kKP + i] = static_cast<int>(e.y);
            p.cand_score[li * kKP + i] = __uint_as_float(e.x);
          }
          if (lane == 0) p.cand_cnt[li] = n;
        }
        if (live) p.list_tau[qrow * p.lists + kParts * chunk + half] = tau;
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR > 1) cluster_sync_all();      // no CTA of the pair leaves (or frees TMEM) while the other can still reach it
  if (warp == 2) {
    tc_fence_after();
    if (PAIR > 1) tmem_dealloc_cg2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
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

// bitonic sort of a power-of-two array by one warp (stages separated by __syncwarp)
__device__ void warp_bitonic_sort(Cand* c, int n_pow2, int lane) {
  for (int k = 2; k <= n_pow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n_pow2; i += 32) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;
          const Cand a = c[i], b = c[l];
          if (up ? cand_before(b, a) : cand_before(a, b)) { c[i] = b; c[l] = a; }
        }
      }
      __syncwarp();
    }
}

// One warp per query (no block-wide barriers): approximate select -> exact fp64 re-score -> sort.
struct CertParams {             // all null / zero: no certificate
  const float* list_tau;        // [nq][lists]
  const float* center;          // [dim] mu of the gallery frame (null: plain unit rows)
  const float* q_err;           // [nq][4]: |dQ_1|, |dQ_rest|, |Q_1|, |Q_rest| of the fp16 query row (d = rounding residual)
  const float* g_stats;         // [4] maxima over the gallery rows of |G_1|, |G_rest|, |dG_1|, |dG_rest| (unscaled)
  float inv_scale;              // approximate scores are in units of `scale`
  int* uncert;                  // [1 + nq]: count, then the queries that could not be certified
};

// SEL = how many of the kept candidates (best by approximate score) are re-scored exactly: kKP for every query, then - only for
// the queries that pass could not certify (`subset` = their list) - a second pass with SEL = kSel2, whose cut-off lies that
// much further below the exact k-th score.  On concentrated embeddings (an untrained backbone: cosines of 0.997 between
// unrelated images, 100 k+ rows) the first pass leaves most queries unproven and the second proves all of them at 1 / 400 of
// the cost of the exact scan of the gallery.
template <int SEL>
__global__ void __launch_bounds__(256, 3) rerank_kernel(const float* __restrict__ q, const double* __restrict__ q_norm,
                                                     const float* __restrict__ g, const double* __restrict__ g_norm, int dim,
                                                     const int* __restrict__ cand_idx, const float* __restrict__ cand_score,
                                                     const int* __restrict__ cand_cnt, int lists, long long nq, int k,
                                                     long long g_index_base, int* __restrict__ out_idx, double* __restrict__ out_score,
                                                     const CertParams cert, const int* __restrict__ subset) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cap = lists * kKP;
  uint8_t* mine = sm + static_cast<size_t>(warp) * (SEL * sizeof(Cand) + static_cast<size_t>(cap) * 8);
  Cand* sel = reinterpret_cast<Cand*>(mine);                                                     // [SEL] exact stage
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(mine + SEL * sizeof(Cand));   // [cap] approximate stage
  const long long slot = 1LL * blockIdx.x * (blockDim.x >> 5) + warp;
  if (slot >= (subset != nullptr ? static_cast<long long>(subset[0]) : nq)) return;
  const long long qi = subset != nullptr ? subset[1 + slot] : slot;
  for (int i = lane; i < SEL; i += 32) { sel[i].score = -INFINITY; sel[i].idx = INT_MAX; sel[i].pad = 0; }
  // unique order-preserving key: (approximate score desc, gallery index asc)  ==  larger key first
  int total = 0;
  for (int c = 0; c < lists; ++c) {
    const int n = cand_cnt[qi * lists + c];
    for (int i = lane; i < n; i += 32) {
      const long long src = (qi * lists + c) * kKP + i;
      keys[total + i] = (static_cast<unsigned long long>(fkey(cand_score[src])) << 32) |
                        static_cast<unsigned long long>(0xffffffffu - static_cast<uint32_t>(cand_idx[src]));
    }
    total += n;
  }
  __syncwarp();
  // level 1: of the survivors of all chunks only the best kKP by approximate (fp16 tensor-core) score can contain
  // the exact top-k (same k + 28 slack argument as inside a chunk).  64-bit radix select, two bits per round; the keys
  // are unique, so it can stop as soon as exactly kKP keys reach the threshold
  unsigned long long thr = 0ULL;
  int cge = total;
#pragma unroll 1
  for (int bit = 62; bit >= 0 && cge > SEL; bit -= 2) {
    const unsigned long long c1 = thr | (1ULL << bit), c2 = thr | (2ULL << bit), c3 = thr | (3ULL << bit);
    int n1 = 0, n2 = 0, n3 = 0;
    for (int i = lane; i < total; i += 32) {
      const unsigned long long key = keys[i];
      n1 += (key >= c1) ? 1 : 0;
      n2 += (key >= c2) ? 1 : 0;
      n3 += (key >= c3) ? 1 : 0;
    }
    n1 = __reduce_add_sync(0xffffffffu, n1);
    n2 = __reduce_add_sync(0xffffffffu, n2);
    n3 = __reduce_add_sync(0xffffffffu, n3);
    if (n3 >= SEL) { thr = c3; cge = n3; }
    else if (n2 >= SEL) { thr = c2; cge = n2; }
    else if (n1 >= SEL) { thr = c1; cge = n1; }
  }
  int nsel = 0;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int i0 = 0; i0 < total; i0 += 32) {
    const int i = i0 + lane;
    const bool keep = i < total && keys[i] >= thr;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    const int pos = nsel + __popc(m & lt_mask);
    if (keep && pos < SEL) sel[pos].idx = static_cast<int>(0xffffffffu - static_cast<uint32_t>(keys[i] & 0xffffffffULL));
    nsel += __popc(m);
  }
  nsel = min(nsel, SEL);
  __syncwarp();
  // level 2: exact fp64 cosine of those <= kKP candidates from the fp32 embeddings (fixed summation order: lane-strided
  // chains, xor tree).  The kernel is bound by the latency of the gathered gallery rows: the norms are fetched up front
  // and four rows (16 x 16 B per lane) are in flight per step
  const float* qr = q + qi * dim;
  const double nqd = fmax(q_norm[qi], 1e-8);
  for (int i = lane; i < nsel; i += 32) sel[i].score = fmax(g_norm[sel[i].idx], 1e-8);     // parked until the dot product lands
  constexpr int kJ = kMaxKB * kBK / 128;
  float4 qa[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    const int d = lane * 4 + j * 128;
    qa[j] = d < dim ? __ldg(reinterpret_cast<const float4*>(qr + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncwarp();
  for (int i = 0; i < nsel; i += 4) {
    float4 b[4][kJ];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float* r = g + 1LL * sel[min(i + c, nsel - 1)].idx * dim;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const int d = lane * 4 + j * 128;
        b[c][j] = d < dim ? __ldg(reinterpret_cast<const float4*>(r + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    double acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double a0 = 0.0;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        if (lane * 4 + j * 128 < dim) {
          a0 = fma(static_cast<double>(qa[j].x), static_cast<double>(b[c][j].x), a0);
          a0 = fma(static_cast<double>(qa[j].y), static_cast<double>(b[c][j].y), a0);
          a0 = fma(static_cast<double>(qa[j].z), static_cast<double>(b[c][j].z), a0);
          a0 = fma(static_cast<double>(qa[j].w), static_cast<double>(b[c][j].w), a0);
        }
      }
      acc[c] = a0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
    if (lane < 4 && i + lane < nsel) {
      const double a = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
      sel[i + lane].score = a / (nqd * sel[i + lane].score);
    }
  }
  __syncwarp();
  warp_bitonic_sort(sel, SEL, lane);
  for (int i = lane; i < k; i += 32) {
    const bool ok = i < nsel;
    out_idx[qi * k + i] = ok ? static_cast<int>(sel[i].idx + g_index_base) : -1;
    out_score[qi * k + i] = ok ? sel[i].score : -INFINITY;
  }
  if (cert.uncert != nullptr) {
    // a_cut: no dropped candidate had an approximate score above it
    float a_cut = -INFINITY;
    for (int c = lane; c < lists; c += 32) a_cut = fmaxf(a_cut, cert.list_tau[qi * lists + c]);
    a_cut = warp_max(a_cut);
    if (total > SEL) a_cut = fmaxf(a_cut, fkey_inv(static_cast<uint32_t>(thr >> 32)));
    // q^ . mu: the constant the centring removed from every score of this query
    double qmu = 0.0;
    if (cert.center != nullptr) {
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const int d = lane * 4 + j * 128;
        if (d < dim) {
          const float4 m = __ldg(reinterpret_cast<const float4*>(cert.center + d));
          qmu = fma(static_cast<double>(qa[j].x), static_cast<double>(m.x), qmu);
          qmu = fma(static_cast<double>(qa[j].y), static_cast<double>(m.y), qmu);
          qmu = fma(static_cast<double>(qa[j].z), static_cast<double>(m.z), qmu);
          qmu = fma(static_cast<double>(qa[j].w), static_cast<double>(m.w), qmu);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) qmu += __shfl_xor_sync(0xffffffffu, qmu, o);
      qmu /= nqd;
    }
    if (lane == 0 && a_cut > -INFINITY) {
      // |Q~ . G~ - Q . G| <= |dQ_1||G_1| + |dQ_r||G_r| + |Q~_1||dG_1| + |Q~_r||dG_r| (first coordinate and the rest bounded
      // separately: the query's first coordinate is O(1), everything else is of the size of the spread), plus the fp32
      // accumulation of 512 products and the fp32 arithmetic that built the rows
      const float* qe = cert.q_err + 4 * qi;
      const double G1 = cert.g_stats[0], Gr = cert.g_stats[1], dG1 = cert.g_stats[2], dGr = cert.g_stats[3];
      const double E = qe[0] * G1 + qe[1] * Gr + qe[2] * dG1 + qe[3] * dGr + 4e-5 * (qe[2] * G1 + qe[3] * Gr);
      const bool proven = nsel >= k && static_cast<double>(a_cut) * cert.inv_scale + E < sel[k - 1].score - qmu;
      if (!proven) cert.uncert[1 + atomicAdd(cert.uncert, 1)] = static_cast<int>(qi);
    }
  }
}

// Exact top-k of the queries the certificate could not prove: one CTA per such query scans the whole gallery with the re-rank's
// own fp64 cosine (same summation order, so a row scores identically on both paths), keeping a pruned candidate buffer.
constexpr int kExBlock = 1024;                    // gallery rows per round
constexpr int kExCap = 2048;                      // candidate buffer (a round's survivors always fit behind the kept k)
__global__ void __launch_bounds__(256) exact_topk_kernel(const float* __restrict__ q, const double* __restrict__ q_norm, const float* __restrict__ g,
                                                         const double* __restrict__ g_norm, long long ng, int dim, int k, long long self_offset,
                                                         int exclude_self, long long g_index_base, const int* __restrict__ uncert,
                                                         int* __restrict__ out_idx, double* __restrict__ out_score) {
  pdl_grid_sync();
  __shared__ Cand buf[kExCap];
  __shared__ int s_cnt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_flag = uncert[0];
  constexpr int kJ = kMaxKB * kBK / 128;
  for (int f = blockIdx.x; f < n_flag; f += gridDim.x) {
    const long long qi = uncert[1 + f];
    const float* qr = q + qi * dim;
    const double nqd = fmax(q_norm[qi], 1e-8);
    float4 qa[kJ];
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const int d = lane * 4 + j * 128;
      qa[j] = d < dim ? __ldg(reinterpret_cast<const float4*>(qr + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const long long self_row = exclude_self ? self_offset + qi : -1;
    if (threadIdx.x == 0) s_cnt = 0;
    Cand cut; cut.score = -INFINITY; cut.idx = INT_MAX; cut.pad = 0;       // keep only rows ranking ahead of `cut` once k are held
    __syncthreads();
    for (long long r0 = 0; r0 < ng; r0 += kExBlock) {
      for (int rr = warp; rr < kExBlock; rr += 8) {
        const long long row = r0 + rr;
        if (row >= ng) break;
        const float* gr = g + row * dim;
        double a0 = 0.0;
#pragma unroll
        for (int j = 0; j < kJ; ++j) {
          const int d = lane * 4 + j * 128;
          if (d < dim) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(gr + d));
            a0 = fma(static_cast<double>(qa[j].x), static_cast<double>(b.x), a0);
            a0 = fma(static_cast<double>(qa[j].y), static_cast<double>(b.y), a0);
            a0 = fma(static_cast<double>(qa[j].z), static_cast<double>(b.z), a0);
            a0 = fma(static_cast<double>(qa[j].w), static_cast<double>(b.w), a0);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        if (lane == 0 && row != self_row) {
          Cand c; c.score = a0 / (nqd * fmax(g_norm[row], 1e-8)); c.idx = static_cast<int>(row); c.pad = 0;
          if (cand_before(c, cut)) buf[atomicAdd(&s_cnt, 1)] = c;
        }
      }
      __syncthreads();
      const int n = s_cnt;
      __syncthreads();                                             // every thread has the SAME n before anyone appends again
      if (n > kExCap - kExBlock || r0 + kExBlock >= ng) {          // CTA-uniform: sort, keep the best k, tighten the cut
        for (int i = n + threadIdx.x; i < kExCap; i += blockDim.x) { buf[i].score = -INFINITY; buf[i].idx = INT_MAX; buf[i].pad = 0; }
        __syncthreads();
        bitonic_sort(buf, kExCap);
        if (threadIdx.x == 0) s_cnt = min(n, k);
        if (n >= k) cut = buf[k - 1];
        __syncthreads();
      }
    }
    const int n = s_cnt;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      out_idx[qi * k + i] = i < n ? static_cast<int>(buf[i].idx + g_index_base) : -1;
      out_score[qi * k + i] = i < n ? buf[i].score : -INFINITY;
    }
    __syncthreads();
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

// fp32 rows -> fp16 rows + fp64 norms (fixed summation order: lane-strided chains, xor tree).  frame == null: plain unit
// rows.  Otherwise frame = [mu (dim) | w (dim) | |mu|] (b200_gallery_frame) and y = H x^ = x^ - 2 w (w . x^):
//   role 1 (gallery): row = scale * [y_1 - |mu|, y_2 ..]            role 0 (query): row = y
// err[row] (nullable) = {|d_1|, |d_rest|, |row_1|, |row_rest|} of the fp16 row (d = fp16 - exact, unscaled); stats (nullable, 4
// floats the caller zeroes): running maxima over the rows of {|row_1|, |row_rest|, |d_1|, |d_rest|}.
__global__ void __launch_bounds__(256) gallery_prepare_kernel(const float* __restrict__ x, const float* __restrict__ frame, int role, float scale,
                                                              __half* __restrict__ out, double* __restrict__ norm, float* __restrict__ err,
                                                              float* __restrict__ stats, long long n, int dim) {
  pdl_grid_sync();
  const long long row = (1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + row * dim;
  const float* mu = frame;
  const float* w = frame != nullptr ? frame + dim : nullptr;
  double ss = 0.0, wx = 0.0;
  for (int d = lane * 4; d < dim; d += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xr + d));
    ss = fma(static_cast<double>(v.x), static_cast<double>(v.x), ss);
    ss = fma(static_cast<double>(v.y), static_cast<double>(v.y), ss);
    ss = fma(static_cast<double>(v.z), static_cast<double>(v.z), ss);
    ss = fma(static_cast<double>(v.w), static_cast<double>(v.w), ss);
    if (w != nullptr) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + d));
      wx = fma(static_cast<double>(v.x), static_cast<double>(ww.x), wx);
      wx = fma(static_cast<double>(v.y), static_cast<double>(ww.y), wx);
      wx = fma(static_cast<double>(v.z), static_cast<double>(ww.z), wx);
      wx = fma(static_cast<double>(v.w), static_cast<double>(ww.w), wx);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); wx += __shfl_xor_sync(0xffffffffu, wx, o); }
  const double nr = sqrt(ss);
  if (lane == 0) norm[row] = nr;
  const double invd = 1.0 / fmax(nr, 1e-8);
  const float inv = static_cast<float>(invd);
  const float two_wx = static_cast<float>(2.0 * wx * invd);                   // 2 (w . x^)
  const double mu_norm = (frame != nullptr && role == 1) ? static_cast<double>(frame[2 * dim]) : 0.0;
  float r1 = 0.f, r2 = 0.f, c1 = 0.f, c2 = 0.f;     // first coordinate: |d|, |value|; the rest: squared norms
  for (int d = lane * 4; d < dim; d += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xr + d));
    float e[4] = {v.x * inv, v.y * inv, v.z * inv, v.w * inv};
    if (w != nullptr) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + d));
      e[0] -= ww.x * two_wx; e[1] -= ww.y * two_wx; e[2] -= ww.z * two_wx; e[3] -= ww.w * two_wx;
      if (d == 0) {       // the first coordinate carries the large numbers: y_1 = x^_1 - 2 w_1 (w . x^) and y_1 - |mu| in fp64
        const double y1 = static_cast<double>(v.x) * invd - 2.0 * static_cast<double>(ww.x) * wx * invd;
        e[0] = static_cast<float>(y1 - mu_norm);
      }
    }
    const __half2 h0 = __floats2half2_rn(e[0] * scale, e[1] * scale), h1 = __floats2half2_rn(e[2] * scale, e[3] * scale);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&h0);
    o.y = *reinterpret_cast<const uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(out + row * dim + d) = o;
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const float back[4] = {f0.x, f0.y, f1.x, f1.y};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float dlt = back[i] - e[i] * scale;
      if (d == 0 && i == 0) { r1 = fabsf(dlt); c1 = fabsf(back[i]); }
      else { r2 = fmaf(dlt, dlt, r2); c2 = fmaf(back[i], back[i], c2); }
    }
  }
  if (err != nullptr || stats != nullptr) {
    r2 = warp_sum(r2);
    c2 = warp_sum(c2);
    r1 = __shfl_sync(0xffffffffu, r1, 0);
    c1 = __shfl_sync(0xffffffffu, c1, 0);
    // a little head-room for the fp32 evaluation of these norms themselves
    const float is = 1.0f / scale;
    const float d1 = r1 * is * 1.01f + 1e-12f, dr = sqrtf(r2) * is * 1.01f + 1e-12f;
    const float v1 = c1 * is * 1.001f + 1e-12f, vr = sqrtf(c2) * is * 1.001f + 1e-12f;
    if (lane == 0) {
      if (err != nullptr) *reinterpret_cast<float4*>(err + 4 * row) = make_float4(d1, dr, v1, vr);
      if (stats != nullptr) {                      // positive floats order like their bit patterns
        unsigned int* st = reinterpret_cast<unsigned int*>(stats);
        atomicMax(st + 0, __float_as_uint(v1)); atomicMax(st + 1, __float_as_uint(vr));
        atomicMax(st + 2, __float_as_uint(d1)); atomicMax(st + 3, __float_as_uint(dr));
      }
    }
  }
}

// frame of a gallery from its mean unit row: [mu | w | |mu|], w = (u - e_1) / |u - e_1| with u = mu / |mu| (H = I - 2 w w^T maps
// u onto the first axis); one CTA, fp64 reductions
__global__ void __launch_bounds__(256) gallery_frame_kernel(const float* __restrict__ mean, int dim, float* __restrict__ frame) {
  pdl_grid_sync();
  __shared__ double red[256];
  double ss = 0.0;
  for (int d = threadIdx.x; d < dim; d += blockDim.x) ss += static_cast<double>(mean[d]) * mean[d];
  red[threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  const double mn = sqrt(red[0]);
  __syncthreads();
  // v = u - e_1;  |v|^2 = 2 - 2 u_1
  const double u1 = mn > 1e-12 ? mean[0] / mn : 1.0;
  const double vn = sqrt(fmax(2.0 - 2.0 * u1, 0.0));
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    frame[d] = mean[d];
    double v = (mn > 1e-12 ? mean[d] / mn : (d == 0 ? 1.0 : 0.0)) - (d == 0 ? 1.0 : 0.0);
    frame[dim + d] = vn > 1e-9 ? static_cast<float>(v / vn) : 0.f;
  }
  if (threadIdx.x == 0) frame[2 * dim] = static_cast<float>(mn);
}

// sum of the unit rows (for the centre mu = mean unit gallery row): one partial row per CTA, fixed order
__global__ void __launch_bounds__(256) unit_row_sum_kernel(const float* __restrict__ x, float* __restrict__ partial, long long n, int dim,
                                                           long long rows_per_block) {
  pdl_grid_sync();
  __shared__ float red[8][kMaxKB * kBK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kJ = kMaxKB * kBK / 128;
  float4 acc[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long r0 = rows_per_block * blockIdx.x, r1 = min(n, r0 + rows_per_block);
  for (long long row = r0 + warp; row < r1; row += 8) {
    const float* xr = x + row * dim;
    float4 v[kJ];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const int d = lane * 4 + j * 128;
      v[j] = d < dim ? __ldg(reinterpret_cast<const float4*>(xr + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
      ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
#pragma unroll
    for (int j = 0; j < kJ; ++j) { acc[j].x += v[j].x * inv; acc[j].y += v[j].y * inv; acc[j].z += v[j].z * inv; acc[j].w += v[j].w * inv; }
  }
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    const int d = lane * 4 + j * 128;
    if (d < dim) *reinterpret_cast<float4*>(&red[warp][d]) = acc[j];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][d];
    partial[1LL * blockIdx.x * dim + d] = a;
  }
}
__global__ void scale_vec_kernel(float* v, int n, float f) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] *= f;
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

// Pair verification scores (engine/controller.py:60-68 with similarity_f of configs/dog_fe/fe_dogs_config.py:89-93):
// out[p] = (cos(emb[i1[p]], emb[i2[p]]) + 1) / 2, fp32 like the reference (cosine_similarity's eps on each norm), one warp
// per pair, fixed summation order (lane-strided chains, xor tree).
__global__ void __launch_bounds__(256) pair_similarity_kernel(const float* __restrict__ emb, const long long* __restrict__ i1,
                                                              const long long* __restrict__ i2, long long n_pairs, int dim, float eps,
                                                              float* __restrict__ out) {
  pdl_grid_sync();
  const long long p = (1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= n_pairs) return;
  const float* a = emb + i1[p] * dim;
  const float* b = emb + i2[p] * dim;
  float ab = 0.f, aa = 0.f, bb = 0.f;
  for (int d = lane * 4; d < dim; d += 128) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a + d)), y = __ldg(reinterpret_cast<const float4*>(b + d));
    ab = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, ab))));
    aa = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, aa))));
    bb = fmaf(y.x, y.x, fmaf(y.y, y.y, fmaf(y.z, y.z, fmaf(y.w, y.w, bb))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ab += __shfl_xor_sync(0xffffffffu, ab, o);
    aa += __shfl_xor_sync(0xffffffffu, aa, o);
    bb += __shfl_xor_sync(0xffffffffu, bb, o);
  }
  if (lane == 0) out[p] = (ab / (fmaxf(sqrtf(aa), eps) * fmaxf(sqrtf(bb), eps)) + 1.0f) * 0.5f;
}

struct Layout {
  long long scratch, cand_idx, cand_score, cand_cnt, list_tau, tau, tau_shared, total;
  int q_blocks, lists;
  long long unc1;              // int [1 + nq]: the queries the first certification pass left open
  long long witness_rows;      // rows [0, witness_rows) seed the thresholds (0 = single pass, unseeded)
  int chunks;                  // gallery chunks of the list pass
  long long chunk_rows;
  int full_groups = 0, tail_chunks = 0;      // ragged last wave (see plan_layout); tail_chunks = 0: off
  long long tail_chunk_rows = 0;
  long long ng;                // gallery rows
  float* tau_ptr;              // [q_blocks*128][2] thresholds handed from the witness pass to the list pass
};

// chunks for `rows` gallery rows.  A chunk starts from the seeded threshold again, so every extra chunk costs ~KP ln(..)
// more candidate appends per query and one more pair of lists to merge: with at least two query blocks per SM there is
// one chunk; below that, enough (q_block, chunk) units to fill the SMs ~3 times, chunks >= 16 tiles, and the unit count
// close to a multiple of the CTA count (the units are equal-sized: a ragged last wave is pure loss)
// CTA pairs (cta_group::2) whenever there are at least two query blocks (B200_GALLERY_PAIR=0: single CTAs, for A/B measurements)
bool use_pair(int q_blocks) {
  static const bool env_off = [] { const char* e = getenv("B200_GALLERY_PAIR"); return e != nullptr && e[0] == '0'; }();
  return !env_off && q_blocks >= 2;
}

// How many chunks the gallery is cut into.  Work units are (query group, chunk) pairs dealt round-robin to `slots` CTAs (or
// CTA pairs).  Few query groups need chunks to fill the machine at all; with many (>= 2 per slot) one chunk is best even when
// the last wave is ragged: 391 query blocks x 125 k rows measured 11.2 ms in one chunk (2.64 waves) against 11.7 ms in three
// (7.93 waves) - every extra chunk restarts the candidate lists from the witness threshold, and that costs more than the tail.
void pick_chunks(long long rows, int q_groups, int slots, int* chunks, long long* chunk_rows) {
  const long long tiles = (rows + kBN - 1) / kBN;
  const long long cap = std::max<long long>(1, std::min<long long>(tiles / 16, 64 / kParts - 1));      // lists = kParts * chunks <= 64
  long long lo = 1, hi = 1;
  if (q_groups < 2 * slots) {
    lo = std::max<long long>(1, std::min<long long>((3LL * slots + q_groups - 1) / q_groups, cap));
    hi = std::min(cap, lo * 2);
  }
  long long best = lo;
  double best_waste = 1e9;
  for (long long c = lo; c <= hi; ++c) {
    const long long units = c * q_groups, ctas = std::min<long long>(units, slots);
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
  L.witness_rows = (ng >= 4LL * kWitnessRows) ? kWitnessRows : 0;      // small galleries: the unseeded flood is cheaper
  const int q_groups = use_pair(L.q_blocks) ? (L.q_blocks + 1) / 2 : L.q_blocks;
  const int slots = use_pair(L.q_blocks) ? sms / 2 : sms;
  pick_chunks(ng, q_groups, slots, &L.chunks, &L.chunk_rows);
  // Ragged last wave.  With one chunk per query group the units are dealt in waves of `slots`; 196 groups on 74 CTA pairs are
  // 2.65 waves, i.e. 26 pairs idle for a whole unit (12 % of the pass).  Only the r = q_groups mod slots groups of the last
  // wave are cut into c gallery chunks - their r * c smaller units take ceil(r c / slots) / c of a unit time instead of 1 -
  // so that the price of chunking (separate candidate lists, restarted thresholds) is paid by those queries alone.
  // B200_GALLERY_TAIL=0 turns it off.
  static const bool tail_on = [] { const char* e = getenv("B200_GALLERY_TAIL"); return e == nullptr || e[0] != '0'; }();
  const long long tiles = (ng + kBN - 1) / kBN;
  const int r = q_groups % slots;
  if (tail_on && L.chunks == 1 && q_groups >= 2 * slots && r > 0 && tiles >= 64) {
    int best_c = 1;
    double best = 1.0;
    for (int c = 2; c <= std::min<long long>(8, 64 / kParts - 1); ++c) {
      const double t = static_cast<double>((1LL * r * c + slots - 1) / slots) / c;
      if (t < best - 0.05) { best = t; best_c = c; }
    }
    if (best_c > 1) {
      const long long tpc = (tiles + best_c - 1) / best_c;
      L.full_groups = q_groups - r;
      L.tail_chunks = static_cast<int>((tiles + tpc - 1) / tpc);
      L.tail_chunk_rows = tpc * kBN;
    }
  }
  L.lists = kParts * std::max(L.chunks, L.tail_chunks);
  long long off = 0;
  auto take = [&](long long bytes) { long long o = off; off = (off + bytes + 255) / 256 * 256; return o; };
  L.scratch = take(1LL * sms * kParts * kBM * kCap * 8);
  L.cand_idx = take(1LL * L.q_blocks * kBM * L.lists * kKP * 4);
  L.cand_score = take(1LL * L.q_blocks * kBM * L.lists * kKP * 4);
  L.cand_cnt = take(1LL * L.q_blocks * kBM * L.lists * 4);
  L.list_tau = take(1LL * L.q_blocks * kBM * L.lists * 4);
  L.tau = take(1LL * L.q_blocks * kBM * kParts * 4);
  L.tau_shared = take(1LL * L.q_blocks * kBM * 4);
  L.unc1 = take(4LL * (1 + L.q_blocks * kBM));
  L.total = off;
  return L;
}

template <int PAIR>
int launch_one(const CUtensorMap& tq, const CUtensorMap& tg, const FilterParams& p, int ctas, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(cosine_filter_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
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
  if (PAIR > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = PAIR; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    n = 2;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  B200_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cosine_filter_kernel<PAIR>, tq, tg, p));
  b200_count_launch();
  return B200_OK;
}

// the passes of one top-k call
int launch_filter(const CUtensorMap& tq, const CUtensorMap& tg, const CUtensorMap& tg_half, FilterParams p, const Layout& L, cudaStream_t st) {
  const int sms = b200_num_sms();
  const bool pair = use_pair(L.q_blocks);
  auto run = [&](int q_blocks, int chunks) -> int {
    const int q_groups = pair ? (q_blocks + 1) / 2 : q_blocks;
    const int units = p.tail_chunks > 0 ? p.full_groups * chunks + (q_groups - p.full_groups) * p.tail_chunks : q_groups * chunks;
    if (pair) {
      FilterParams pp = p;
      pp.idesc = gemm::make_idesc(false, kBN, false, 256);
      return launch_one<2>(tq, tg_half, pp, 2 * std::min(units, sms / 2), st);
    }
    return launch_one<1>(tq, tg, p, std::min(units, sms), st);
  };
  int rc;
  p.g_begin = 0;
  if (L.witness_rows > 0) {
    p.ng = L.witness_rows; p.chunks = 1; p.chunk_rows = L.witness_rows;
    p.witness = 1; p.tau_init = nullptr; p.tau_out = L.tau_ptr;
    p.full_groups = 0; p.tail_chunks = 0; p.tail_chunk_rows = 0;
    if ((rc = run(L.q_blocks, 1))) return rc;
  }
  p.ng = L.ng; p.chunks = L.chunks; p.chunk_rows = L.chunk_rows;
  p.full_groups = L.full_groups; p.tail_chunks = L.tail_chunks; p.tail_chunk_rows = L.tail_chunk_rows;
  p.witness = 0; p.tau_init = L.witness_rows ? L.tau_ptr : nullptr; p.tau_out = nullptr;
  return run(L.q_blocks, L.chunks);
}

}  // namespace

extern "C" int b200_gallery_prepare_ex(const float* emb, const float* frame, int role, float scale, void* f16_rows, double* norm, float* err,
                                       float* stats, long long n, int dim, void* stream) {
  B200_REQUIRE(dim % 4 == 0, "gallery_prepare: dim must be a multiple of 4");
  B200_REQUIRE(scale > 0.f && (role == 0 || role == 1), "gallery_prepare: scale must be positive, role 0 (query) or 1 (gallery)");
  if (n == 0) return B200_OK;
  const long long blocks = (n * 32 + 255) / 256;
  launch_pdl(gallery_prepare_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
             emb, frame, role, scale, reinterpret_cast<__half*>(f16_rows), norm, err, stats, n, dim);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_gallery_prepare(const float* emb, void* unit_f16, double* norm, long long n, int dim, void* stream) {
  return b200_gallery_prepare_ex(emb, nullptr, 0, 1.0f, unit_f16, norm, nullptr, nullptr, n, dim, stream);
}

// frame [2 * dim + 1] floats from the mean unit gallery row (b200_unit_row_mean)
extern "C" int b200_gallery_frame(const float* mean, int dim, float* frame, void* stream) {
  B200_REQUIRE(dim % 4 == 0 && dim > 0, "gallery_frame: dim must be a positive multiple of 4");
  launch_pdl(gallery_frame_kernel, dim3(1), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), mean, dim, frame);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_unit_row_mean_blocks(long long n) {
  const long long b = (n + 2047) / 2048;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(b, 4LL * b200_num_sms())));
}

// mean[d] = (1 / n) sum_rows x[row][d] / |x[row]|: the centre b200_gallery_prepare_ex subtracts.  partial: fp32
// [b200_unit_row_mean_blocks(n)][dim] scratch.
extern "C" int b200_unit_row_mean(const float* emb, long long n, int dim, float* mean, float* partial, void* stream) {
  B200_REQUIRE(dim % 4 == 0 && dim <= kMaxKB * kBK && n > 0, "unit_row_mean: dim must be a multiple of 4, <= 512, n > 0");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = b200_unit_row_mean_blocks(n);
  const long long rpb = (n + blocks - 1) / blocks;
  launch_pdl(unit_row_sum_kernel, dim3(blocks), dim3(256), 0, st, emb, partial, n, dim, rpb);
  B200_LAUNCH_CHECK();
  int rc = splitk_reduce(partial, mean, dim, blocks, 0, st, dim);
  if (rc) return rc;
  launch_pdl(scale_vec_kernel, dim3((dim + 255) / 256), dim3(256), 0, st, mean, dim, static_cast<float>(1.0 / static_cast<double>(n)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" long long b200_cosine_topk_workspace_bytes(long long nq, long long ng, int dim, int k) {
  (void)dim; (void)k;
  if (nq <= 0 || ng <= 0) return 256;
  return plan_layout(nq, ng).total;
}

static int cosine_topk_impl(const float* q, const void* q_unit_f16, const double* q_norm, long long nq, const float* g,
                            const void* g_unit_f16, const double* g_norm, long long ng, int dim, int k,
                            long long exclude_self_offset, long long g_index_base, int* out_idx, double* out_score,
                            void* workspace, long long workspace_bytes, CertParams cert, void* stream) {
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
  p.wait_ns = static_cast<uint32_t>(b200_wait_ns());
  p.scratch = reinterpret_cast<uint2*>(ws + L.scratch);
  p.cand_idx = reinterpret_cast<int*>(ws + L.cand_idx);
  p.cand_score = reinterpret_cast<float*>(ws + L.cand_score);
  p.cand_cnt = reinterpret_cast<int*>(ws + L.cand_cnt);
  p.list_tau = reinterpret_cast<float*>(ws + L.list_tau);
  cert.list_tau = p.list_tau;
  L.ng = ng;
  L.tau_ptr = reinterpret_cast<float*>(ws + L.tau);
  p.tau_shared = reinterpret_cast<uint32_t*>(ws + L.tau_shared);
  B200_CHECK_CUDA(cudaMemsetAsync(p.tau_shared, 0, sizeof(uint32_t) * L.q_blocks * kBM, st));
  if (L.tail_chunks > 0) {
    // queries of the full waves fill only their first kParts * chunks lists: the others must read as empty, and their
    // thresholds as "nothing dropped" (0xff.. = NaN, which fmaxf ignores)
    B200_CHECK_CUDA(cudaMemsetAsync(p.cand_cnt, 0, sizeof(int) * L.q_blocks * kBM * L.lists, st));
    B200_CHECK_CUDA(cudaMemsetAsync(p.list_tau, 0xff, sizeof(float) * L.q_blocks * kBM * L.lists, st));
  }
  CUtensorMap tq, tg;
  int rc = gemm::encode_tmap_2d(&tq, false, q_unit_f16, dim, nq, dim, kBK, kBM);
  if (rc) return rc;
  rc = gemm::encode_tmap_2d(&tg, false, g_unit_f16, dim, ng, dim, kBK, kBN);
  if (rc) return rc;
  CUtensorMap tg_half;       // one CTA's half of a gallery tile in pair mode
  rc = gemm::encode_tmap_2d(&tg_half, false, g_unit_f16, dim, ng, dim, kBK, kBN / 2);
  if (rc) return rc;
  rc = launch_filter(tq, tg, tg_half, p, L, st);
  if (rc) return rc;
  auto launch_rerank = [&](auto kernel, int sel, const CertParams& c, const int* subset) -> int {
    const int per_warp = sel * static_cast<int>(sizeof(Cand)) + L.lists * kKP * 8;      // exact stage + candidate keys of one query
    B200_REQUIRE(per_warp <= 200 * 1024, "cosine_topk: too many candidate lists");
    const int wpc = std::max(1, std::min(8, (200 * 1024) / per_warp));                  // queries (warps) per CTA
    const int smem2 = wpc * per_warp;
    if (smem2 > 48 * 1024) B200_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    launch_pdl(kernel, dim3(static_cast<unsigned>((nq + wpc - 1) / wpc)), dim3(32 * wpc), smem2, st, q, q_norm, g, g_norm, dim, p.cand_idx,
               p.cand_score, p.cand_cnt, L.lists, nq, k, g_index_base, out_idx, out_score, c, subset);
    B200_LAUNCH_CHECK();
    return B200_OK;
  };
  if (cert.uncert == nullptr) return launch_rerank(rerank_kernel<kKP>, kKP, cert, nullptr);
  // certified: pass 1 leaves its open queries in the workspace list, pass 2 re-scores kSel2 candidates for those and leaves what
  // it cannot prove either in the caller's list, which the exact scan of the gallery then serves (normally nobody)
  int* unc1 = reinterpret_cast<int*>(ws + L.unc1);
  B200_CHECK_CUDA(cudaMemsetAsync(unc1, 0, sizeof(int), st));
  B200_CHECK_CUDA(cudaMemsetAsync(cert.uncert, 0, sizeof(int), st));
  CertParams c1 = cert;
  c1.uncert = unc1;
  if ((rc = launch_rerank(rerank_kernel<kKP>, kKP, c1, nullptr))) return rc;
  if ((rc = launch_rerank(rerank_kernel<kSel2>, kSel2, cert, unc1))) return rc;
  launch_pdl(exact_topk_kernel, dim3(static_cast<unsigned>(std::min<long long>(nq, 4LL * b200_num_sms()))), dim3(256), 0, st, q, q_norm, g, g_norm,
             ng, dim, k, p.self_offset, p.exclude_self, g_index_base, cert.uncert, out_idx, out_score);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200_cosine_topk(const float* q, const void* q_unit_f16, const double* q_norm, long long nq, const float* g,
                                const void* g_unit_f16, const double* g_norm, long long ng, int dim, int k,
                                long long exclude_self_offset, long long g_index_base, int* out_idx, double* out_score,
                                void* workspace, long long workspace_bytes, void* stream) {
  return cosine_topk_impl(q, q_unit_f16, q_norm, nq, g, g_unit_f16, g_norm, ng, dim, k, exclude_self_offset, g_index_base, out_idx, out_score,
                          workspace, workspace_bytes, CertParams{}, stream);
}

// b200_cosine_topk with the exactness certificate.  q_f16 / g_f16 = rows of b200_gallery_prepare_ex in the SAME frame (roles 0 /
// 1; frame may be null for both = plain unit rows), q_err = the queries' err rows, g_stats = the maxima the gallery call
// accumulated.  uncertified: int [1 + nq], receives the number of queries re-done by the exact scan and their indices.
extern "C" int b200_cosine_topk_certified(const float* q, const void* q_f16, const double* q_norm, const float* q_err, long long nq,
                                          const float* g, const void* g_f16, const double* g_norm, const float* frame, float g_scale,
                                          const float* g_stats, long long ng, int dim, int k, long long exclude_self_offset,
                                          long long g_index_base, int* out_idx, double* out_score, int* uncertified, void* workspace,
                                          long long workspace_bytes, void* stream) {
  B200_REQUIRE(q_err != nullptr && g_stats != nullptr && uncertified != nullptr && g_scale > 0.f, "cosine_topk_certified: missing certificate inputs");
  CertParams cert{};
  cert.center = frame; cert.q_err = q_err; cert.g_stats = g_stats; cert.inv_scale = 1.0f / g_scale; cert.uncert = uncertified;
  return cosine_topk_impl(q, q_f16, q_norm, nq, g, g_f16, g_norm, ng, dim, k, exclude_self_offset, g_index_base, out_idx, out_score,
                          workspace, workspace_bytes, cert, stream);
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

extern "C" int b200_pair_similarity(const float* emb, long long n, int dim, const long long* i1, const long long* i2, long long n_pairs,
                                    float* out, void* stream) {
  B200_REQUIRE(dim % 4 == 0 && dim > 0 && n >= 0, "pair_similarity: dim must be a positive multiple of 4 (got %d)", dim);
  if (n_pairs == 0) return B200_OK;
  launch_pdl(pair_similarity_kernel, dim3(static_cast<unsigned>((n_pairs * 32 + 255) / 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
             emb, i1, i2, n_pairs, dim, 1e-8f, out);
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
