// Batched, bucket-stationary MIH (K2-K5 for a BATCH of queries): every probed bucket is read from HBM once
// per radius level for all the queries that probe it, instead of once per query.
//
// Same algorithm and the same answers as mih_search_kernel (mih.cuh) / SearchWorker::find
// (src/search_worker.cc:65-264); what changes is the loop order.  The reference walks query -> radius ->
// table -> probe -> bucket members.  Here, per radius level r (all still-active queries together):
//   1. bmih_probe_kernel   every (query, table, probe) of the level is unranked (enumerate_entry,
//                          search_worker.cc:230-264); non-empty buckets are counted, then - second pass -
//                          the query is written into its bucket's list (a counting sort by bucket);
//   2. bmih_items_kernel   each probed bucket is cut into work items of <= kBmihCPI codes x <= kBmihQT queries;
//   3. bmih_verify_kernel  persistent CTAs pull items; a CTA stages the item's queries in shared memory,
//                          streams the bucket's codes through registers with 128-bit loads and tests every
//                          code against every query (XOR + POPC + MIN, optional half-POPC prefilter) - the
//                          full-code verification of search_worker.cc:249-257.  Survivors of the distance
//                          threshold pass the first-discoverer de-duplication (knn_found_, :183-190, decided
//                          from the code itself) and are appended to the query's global candidate buffer; the
//                          threshold tau[q] is the exact running k-th distance (global histogram + atomicMin);
//   4. bmih_settle_kernel  per query: sort the buffer, keep the k best (the heap of :192-197), apply the
//                          strict m-aware stop rule (:201-207) and build the next active list.
// Before level 0 a bootstrap kernel bounds every query's k-th distance from a sample of its own buckets, so
// that the candidate buffers stay small from the first level on.  A query whose candidate buffer overflows
// all the same (adversarial ties) is re-run by the per-query kernel, so the result is exact whatever the data.
#pragma once
#include <math.h>
#include "mih.cuh"
#include "scan.cuh"
#include "xchg.cuh"

namespace vc {

#ifndef VC_BMIH_THREADS
#define VC_BMIH_THREADS 256
#endif
#ifndef VC_U4
#define VC_U4 4
#endif
constexpr int kBmihThreads = VC_BMIH_THREADS;
constexpr int kBmihQT = 32;          // queries per work item
constexpr int kProbeMasks = 2048;    // probe masks of one table and step the probe kernel keeps in shared memory (C(16, 3) = 560; beyond: unranked per probe)
constexpr int kBmihSort = 4096;      // entries the settle kernel sorts at a time (shared memory); k must stay below half of it
constexpr int kBmihCapMin = 4096;    // candidate-buffer entries per query: at least this, see bmih_cap_for
constexpr int kBmihHitQ = 32 * kBmihQT + 32;   // deferred hits per warp of the verify kernel (bv_hitq): what one warp step can add, plus an undrained rest
constexpr int kBmihU4 = VC_U4;           // 128-bit loads per thread per step (short-bucket variant)
#ifndef VC_VERIFY_CTAS
#define VC_VERIFY_CTAS 4   // 64-bit codes: 4 CTAs (32 warps, 64 registers) per SM - the kernel is latency-bound, 15.5 instead of 16.7 ms per search
#endif
#ifndef VC_VERIFY_CTAS_W2
#define VC_VERIFY_CTAS_W2 4   // 128- and 256-bit codes: 4 CTAs per SM (64 registers) measured 5 % faster than 3 (80) on C3 and C5 (profiles/ab_r02.md)
#endif
#ifndef VC_KEY_SUBST
#define VC_KEY_SUBST 1       // bucket key substituted into the staged queries (a sharper one-POPC-per-64-bit filter), see bmih_verify_kernel
#endif
#ifndef VC_HIT_QUEUE
#define VC_HIT_QUEUE 1       // codes that pass the filter are queued per warp and re-checked 32 at a time instead of inside the distance loop
#endif
#ifndef VC_PF_DIST
#define VC_PF_DIST 4
#endif
#ifndef VC_TAU_EVERY
#define VC_TAU_EVERY 1
#endif
constexpr int kPfDist = VC_PF_DIST;  // L2 prefetch distance of the verify kernel, in warp steps of 2 KB

template <int W, int U4 = kBmihU4> struct BmihCfg {
  static constexpr int C = 2 * U4 / W;                      // codes per thread per step (U4 = 4: 8 / 4 / 2)
  static constexpr int STEP = 2048 / W;                     // unit of the work-item length (16 KB of codes; a multiple of every warp step)
  static constexpr int QS = (2 * W + 1 + 3) / 4 * 4;        // u32 per staged query: words, tau, pad
  static constexpr int HB = 64 * W + 32;
};

// one work item: codes [c0, c1) of table t (bucket order) against queries qlist[qbeg, qbeg + qn)
struct BmihItem { uint32_t t, c0, c1, qbeg, qn; };

struct BmihParams {
  const uint32_t* queries;      // [nq][2W]
  uint32_t nq, k, m, sbits;
  uint32_t radius;              // (highest) radius of the step being processed
  uint32_t r_lo;                // lowest radius of the step: radii [r_lo, radius] are probed together (normally r_lo == radius)
  uint32_t t_begin, t_end;      // ... and its tables [t_begin, t_end): a whole radius (0, m) or one table of it
  uint32_t cpi;                 // codes per work item (multiple of the step size)
  uint32_t cap;                 // candidate-buffer entries per query (bmih_cap_for(k))
  uint32_t pf_tau;              // thresholds (less the probing radius) >= this count as loose: the lower-bound filter would pass too much (bmih_pf_tau)
  uint32_t boot_sample;         // codes per query the threshold bootstrap looks at (0: the default)
  uint32_t spec_tau;            // speculative bound of every query's k-th distance (kInfDist: none), see bmih_decide_kernel
  uint32_t* spec_fail;          // [1] queries whose k-th distance turned out to lie above it (they are redone)
  uint32_t count_in_write;      // items kernel: the writing pass is the only pass, it also keeps the statistics
  uint32_t qt;                  // queries per work item (kBmihQT for the POPC kernel, up to 256 for the tensor-core kernel)
  uint32_t cpi_alt, qt_alt;     // counting pass only: the item geometry of the other verify kernel ...
  uint32_t* n_items_alt;        // ... and its item count, so that the host can choose between the two after one pass
  long long* tc_trace;          // debug (knob tc.trace): per-unit / per-tile clock64 timestamps of CTA 0, [2][256][4]
  unsigned long long* tc_stats; // tensor-core kernel: [0] warp x (tile, query group) units, [1] units with a flagged row, [2] hits appended
  int max_radius;
  const TableDev* tables;       // [m]
  // per-level probe structures
  const uint32_t* active;       // [n_active] query indices
  uint32_t n_active;
  uint32_t* bcount;             // [m << sbits] queries per bucket (pass 0), then cursor (pass 1)
  uint32_t* boffs;              // [(m << sbits) + 1] exclusive scan of bcount
  uint8_t* bflag;               // [m << sbits] or null: 1 = the bucket is probed at radius 0 by some query of this step (its work items go first)
  uint32_t* qlist;              // [total probes] query index per (bucket, slot)
  BmihItem* items;
  uint32_t* n_items;            // [1]
  unsigned long long* bucket_codes;   // [1] codes of all distinct probed buckets, summed over the steps (traffic accounting)
  unsigned long long* pair_count;     // [1] code-query tests = members of all probed buckets over all queries, summed over the steps
  unsigned long long* exec_pairs;     // [1] code-query tests the verify kernel really executed (less: queries leave buckets at their k-th id)
  uint32_t* item_cursor;        // [1]
  // per-query state
  uint64_t* gbuf;               // [nq][cap]
  uint32_t* gcnt;               // [nq]
  uint64_t* gtaukey;            // [nq]
  uint64_t* gglobkey;           // [nq] id-sharded search: a key that at least k codes of the WHOLE database are below (or kEmptyKey)
  uint32_t* gtau;               // [nq]
  uint32_t* ghist;              // [nq][HB]
  uint32_t* gflag;              // [nq] bit0 = buffer overflowed (redo with the per-query kernel), bit1 = finished
  uint32_t* gradius;            // [nq] last radius searched
  unsigned long long* gprobes;  // [nq]
  unsigned long long* gcands;   // [nq]
  uint32_t* next_active;        // [nq]
  uint32_t* n_next;             // [1]
  // brute-force scan through the same machinery: one pseudo table (the main code array in id order, ids = first_id +
  // position), one "bucket" = the whole shard, every query in its list; no de-duplication needed
  uint32_t scan_mode, first_id, id_stride;
};

// ---- 1. probes of one level -------------------------------------------------------------------------------
// pass 0: count queries per non-empty bucket (+ statistics); pass 1: write the query into its bucket's list
template <int W>
__global__ void bmih_probe_kernel(const BmihParams p, int pass) {
  uint32_t per_table = 0;                                     // probes per table: all masks of weight r_lo .. radius
  for (uint32_t r = p.r_lo; r <= p.radius; ++r) per_table += c_binom[p.sbits][r];
  const uint32_t per_q = per_table * (p.t_end - p.t_begin);   // the host keeps a step below 2^32 probes: 32-bit divisions below
  const uint64_t total = (uint64_t)per_q * p.n_active;
  const uint32_t lane = threadIdx.x & 31;
  // the step's masks, unranked once per CTA: unrank_mask walks the binomial table in constant memory with a different index
  // in every lane (serialised), and a CTA unranks the same per_table masks for every query it meets
  __shared__ uint32_t s_mask[kProbeMasks];
  const bool tabled = per_table <= (uint32_t)kProbeMasks;
  if (tabled) {
    for (uint32_t i = threadIdx.x; i < per_table; i += blockDim.x) {
      uint32_t pidx = i, rad = p.r_lo;
      while (pidx >= c_binom[p.sbits][rad]) { pidx -= c_binom[p.sbits][rad]; ++rad; }
      s_mask[i] = unrank_mask(p.sbits, rad, pidx);
    }
    __syncthreads();
  }
  unsigned long long warp_pairs = 0;                          // lane 0: members of all buckets this warp probed
  // whole warps walk the probe list (the statistics below are aggregated per warp: one atomic per distinct query and
  // warp instead of one per probe - the probes of a query are neighbours, and 11 M atomics on one counter cost 3 ms)
  for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; base < total; base += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t it = base + lane;
    const bool valid = it < total;
    uint32_t q = 0xFFFFFFFFu, len = 0, b = 0, key = 0, qkey_of_probe = 1;
    if (valid) {
      const uint32_t a = (uint32_t)it / per_q;
      const uint32_t rem = (uint32_t)it - a * per_q;
      const uint32_t tt = rem / per_table;
      const uint32_t t = p.t_begin + tt;
      uint32_t pidx = rem - tt * per_table, mask;
      if (tabled) mask = s_mask[pidx];
      else {
        uint32_t rad = p.r_lo;
        while (pidx >= c_binom[p.sbits][rad]) { pidx -= c_binom[p.sbits][rad]; ++rad; }
        mask = unrank_mask(p.sbits, rad, pidx);
      }
      q = p.active[a];
      const uint32_t qkey = substring<W>(p.queries + (size_t)q * 2 * W, t, p.sbits);
      key = qkey ^ mask;
      qkey_of_probe = qkey;
      const uint32_t* rp = p.tables[t].row_ptr;
      len = rp[key + 1] - rp[key];
      b = (t << p.sbits) + key;
    }
    if (pass == 0) {
      if (len) atomicAdd(&p.bcount[b], 1u);
      if (len && p.bflag && key == qkey_of_probe) p.bflag[b] = 1;
      if (__any_sync(0xffffffffu, len >= (1u << 26))) {         // giant buckets (degenerate data): 32-bit warp sums could wrap
        if (len) { atomicAdd(&p.gcands[q], (unsigned long long)len); atomicAdd(p.pair_count, (unsigned long long)len); }
      } else {
        const uint32_t peers = __match_any_sync(0xffffffffu, q);
        const uint32_t sum = __reduce_add_sync(peers, len);
        if (valid && sum && lane == (uint32_t)__ffs(peers) - 1) atomicAdd(&p.gcands[q], (unsigned long long)sum);
        warp_pairs += __reduce_add_sync(0xffffffffu, len);
      }
    } else if (len) {
      const uint32_t slot = atomicAdd(&p.bcount[b], 1u);
      p.qlist[p.boffs[b] + slot] = q;
    }
  }
  if (pass == 0 && lane == 0 && warp_pairs) atomicAdd(p.pair_count, warp_pairs);
}

// longest bucket of the dense tables (once per built index): bounds the work items per probed bucket
__global__ void bmih_maxlen_kernel(const TableDev* tables, uint32_t m, uint32_t sbits, uint32_t* out) {
  const uint32_t n_buckets = m << sbits;
  uint32_t mx = 0;
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n_buckets; b += gridDim.x * blockDim.x) {
    const uint32_t* rp = tables[b >> sbits].row_ptr;
    const uint32_t key = b & ((1u << sbits) - 1);
    mx = max(mx, rp[key + 1] - rp[key]);
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(out, mx);
}

// ---- 2. work items -------------------------------------------------------------------------------------------
// write = 0: only count the items (n_items); write = 1: emit descriptors at atomically claimed positions
// phase (with p.bflag): 0 = only the buckets probed at radius 0, 1 = only the others, 2 = all.  A step that starts at radius 0 emits
// its items in two launches, 0 then 1: the queries' own buckets hold their nearest candidates, and with those verified first the
// thresholds have tightened before the bulk of the step (the radius-1 buckets) is streamed.
template <int W>
__global__ void bmih_items_kernel(const BmihParams p, int write, int phase = 2) {
  const uint32_t n_buckets = p.m << p.sbits;
  const uint32_t lane = threadIdx.x & 31;
  // whole warps walk the bucket list: item positions are claimed with one atomic per warp (prefix sum over the lanes),
  // the statistics are summed per warp (170 K atomics on one counter made this kernel 80 us instead of ~15)
  for (uint32_t b0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; b0 < n_buckets; b0 += gridDim.x * blockDim.x) {
    const uint32_t b = b0 + lane;
    uint32_t cnt = 0, len = 0, t = 0, key = 0, nc = 0, nqc = 0, alt = 0;
    const uint32_t* rp = nullptr;
    if (b < n_buckets) cnt = p.boffs[b + 1] - p.boffs[b];
    if (cnt && phase != 2 && (p.bflag[b] != 0) != (phase == 0)) cnt = 0;
    if (cnt) {
      t = b >> p.sbits; key = b & ((1u << p.sbits) - 1);
      rp = p.tables[t].row_ptr;
      len = rp[key + 1] - rp[key];
      nc = (len + p.cpi - 1) / p.cpi; nqc = (cnt + p.qt - 1) / p.qt;
      if (!write && p.n_items_alt) alt = ((len + p.cpi_alt - 1) / p.cpi_alt) * ((cnt + p.qt_alt - 1) / p.qt_alt);
    }
    const uint32_t mine = nc * nqc;
    uint32_t incl = mine;                                    // inclusive prefix sum over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += v; }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (!total) continue;
    uint32_t wbase = 0;
    if (lane == 0) wbase = atomicAdd(p.n_items, total);
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    const uint32_t base = wbase + incl - mine;
    if (!write || p.count_in_write) {
      if (__any_sync(0xffffffffu, len >= (1u << 26))) {      // giant buckets (degenerate data): 32-bit warp sums could wrap
        if (len) atomicAdd(p.bucket_codes, (unsigned long long)len);
      } else {
        const uint32_t wl = __reduce_add_sync(0xffffffffu, len);
        if (lane == 0) atomicAdd(p.bucket_codes, (unsigned long long)wl);
      }
      if (p.n_items_alt) {
        const uint32_t wa = __reduce_add_sync(0xffffffffu, alt);
        if (lane == 0 && wa) atomicAdd(p.n_items_alt, wa);
      }
    }
    if (write && cnt) {
      const uint32_t start = rp[key];
      for (uint32_t c = 0; c < nc; ++c)
        for (uint32_t qc = 0; qc < nqc; ++qc) {
          // the bucket's query list is cut into equal chunks of at most p.qt queries
          const uint32_t qlo = (uint32_t)(((uint64_t)cnt * qc) / nqc), qhi = (uint32_t)(((uint64_t)cnt * (qc + 1)) / nqc);
          BmihItem it;
          it.t = t; it.c0 = start + c * p.cpi; it.c1 = min(start + len, it.c0 + p.cpi);
          it.qbeg = p.boffs[b] + qlo; it.qn = qhi - qlo;
          p.items[base + c * nqc + qc] = it;
        }
    }
  }
}

// ---- 3. verification ---------------------------------------------------------------------------------------
// candidate-buffer entries per query for a search with this k.  A step that looks at C candidates after S were known appends
// about k ln(C / S) of them before the thresholds have settled (each new code is among the k best so far with probability
// k / seen), so the buffer scales with k: 4096 entries for k <= 512, 8 k rounded up to a power of two above that.
__host__ __device__ inline uint32_t bmih_cap_for(uint32_t k) {
  uint32_t cap = kBmihCapMin;
  while (cap < 8u * k) cap <<= 1;
  return cap;
}

// The lower-bound filter popc(x_lo | x_hi) costs half the POPCs of the exact distance but passes more codes: on random bits it is
// ~ N(mean, var) with 0.75 per OR-ed position and 0.5 per position whose partner is the substituted substring (VC_KEY_SUBST).
// Every pass costs the hit path ~ 10 clocks, the exact filter 8 more POPCs = 64 clocks per warp-record of 256 tests: the lower
// bound pays while fewer than ~ 6 of 256 pass, i.e. below mean - 2.2 sigma.  A search step whose queries are mostly above that
// (thresholds fresh from the bootstrap: the first step for 64-bit codes, radii 0 - 2 for 128-bit codes at 125 M codes) runs the
// kernel variant that filters on the exact distance; bmih_decide_kernel counts the loose queries for the host's choice.  (Decided
// per step, not per record: a warp-uniform branch per record in the distance loop cost 4 - 8 % in the steps that do not need it.)
__host__ inline uint32_t bmih_pf_tau(uint32_t W, uint32_t sbits, bool scan_mode) {
  const double s = scan_mode || !VC_KEY_SUBST ? 0.0 : (double)sbits;
  const double dbl = 32.0 * W - s, mean = 0.75 * dbl + 0.5 * s, var = 0.1875 * dbl + 0.25 * s;
  const double t = mean - 2.2 * sqrt(var);
  return t < 1.0 ? 1u : (uint32_t)t;
}

// rare path: a code whose exact distance d passed the staged threshold tau_s of query qid.
// ghist[q][b] holds CUMULATIVE counts: distinct known codes at distance <= b (maintained for b below the
// threshold only - bins at or above it are never consulted again, since thresholds only fall).
template <int W>
__device__ __forceinline__ void bmih_append_impl(const BmihParams* pp, uint32_t qid, uint32_t t, uint32_t d, uint32_t j,
                                                 CodeRegs<W> c, uint32_t* qrec /* shared: query words, then tau */, uint32_t tau_s,
                                                 const uint32_t* qorig = nullptr /* the query's own words when the staged ones were altered */,
                                                 uint32_t r_sub = 0 /* what the staged threshold is short of the real one */) {
  const BmihParams& p = *pp;
  const uint32_t* qo = qorig ? qorig : qrec;
  // first-discoverer test: table t holds this code at substring distance r_own from the query (the radius at
  // which this bucket was probed for this query); it is emitted here only if no other table holds it at a
  // smaller substring distance, or at the same one with a lower table id
  if (!p.scan_mode) {
    uint32_t x[2 * W];
#pragma unroll
    for (int i = 0; i < 2 * W; ++i) x[i] = c.w[i] ^ qo[i];
    const uint32_t r_own = __popc(substring<W>(x, t, p.sbits));
    for (uint32_t t2 = 0; t2 < p.m; ++t2) {
      if (t2 == t) continue;
      const uint32_t sd = __popc(substring<W>(x, t2, p.sbits));
      if (sd < r_own || (sd == r_own && t2 < t)) return;
    }
  }
  const uint64_t key = pack_key(d, p.scan_mode ? p.first_id + j * p.id_stride : p.tables[t].ids[j]);
  if (key >= __ldcg(&p.gtaukey[qid])) return;
  constexpr int HB = BmihCfg<W>::HB;
  uint32_t* gc = p.ghist + (size_t)qid * HB;
  const uint32_t lim = min(tau_s, (uint32_t)(64 * W + 1));      // bins d .. lim-1 count this code
  const uint32_t slot = atomicAdd(&p.gcnt[qid], 1u);
  if (slot < p.cap) p.gbuf[(size_t)qid * p.cap + slot] = key;
  else atomicOr(&p.gflag[qid], 1u);
  if (d >= lim) return;
  for (uint32_t b0 = d; b0 + 1 < lim; ++b0) atomicAdd(&gc[b0], 1u);              // fire and forget
  const uint32_t old = atomicAdd(&gc[lim - 1], 1u);
  if (old + 1 >= p.k) {                                        // k codes are now strictly inside the threshold
    uint32_t nt = lim - 1;
    while (nt > 0 && __ldcg(&gc[nt - 1]) >= p.k) --nt;
    atomicMin(&p.gtau[qid], nt);
    atomicMin(&qrec[2 * W], nt >= r_sub ? nt - r_sub : 0u);
  }
}
template <int W>
__device__ __noinline__ void bmih_append(const BmihParams* pp, uint32_t qid, uint32_t t, uint32_t d, uint32_t j,
                                         CodeRegs<W> c, uint32_t* qrec, uint32_t tau_s) {
  bmih_append_impl<W>(pp, qid, t, d, j, c, qrec, tau_s);
}

// The verify kernel's per-warp state lives at namespace scope, in ONE record per warp, so that (a) the rare paths find the
// warp's slice by themselves (a pointer argument would have to be kept alive inside the distance loop, which has no registers
// to spare) and (b) the distance loop addresses all of it - staged queries, hit queue - from a single 32-bit shared-memory
// address plus constant offsets (ld.shared / atom.shared with explicit addresses below: the generic-pointer form made the
// compiler carry four addresses through the loop and reload them from local memory for every record).
constexpr int kBmihQSMax = 12;                                      // BmihCfg<4>::QS
struct __align__(16) BvWarp {
  uint32_t qrec[kBmihQT * kBmihQSMax];   // staged queries: QS words each - the query words, tau, (substituted substring's distance), pad
  uint32_t qid[kBmihQT];                 // their query indices
  uint32_t cut[kBmihQT];                 // their k-th ids (id cut of the last step of a search), or ~0
  uint32_t hitn, pad_[3];                // entries in hitq
  uint8_t perm[kBmihQT];                 // scratch of the id cut: old staged slot -> new staged slot
  uint16_t hitq[kBmihHitQ];              // deferred hits: warp step of the item << 10 | lane << 5 | staged query slot
};
constexpr uint32_t kBvQid = kBmihQT * kBmihQSMax * 4, kBvCut = kBvQid + kBmihQT * 4, kBvHitN = kBvCut + kBmihQT * 4, kBvHitQ = kBvHitN + 16 + kBmihQT;
static_assert(sizeof(BvWarp) == kBvHitQ + kBmihHitQ * 2 && sizeof(BvWarp) % 16 == 0, "BvWarp layout");
__shared__ BvWarp bv_warp[kBmihThreads / 32];
__shared__ unsigned long long bv_pairs[kBmihThreads / 32];          // tests executed by the warp (statistics)
__shared__ const uint64_t* bv_codes[kMaxTables];                    // table payload pointers, fetched once per CTA
__shared__ const uint32_t* bv_ids[kMaxTables];

__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_add_u32(uint32_t a, uint32_t v) {
  uint32_t o;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory");
  return o;
}
__device__ __forceinline__ void atoms_min_u32(uint32_t a, uint32_t v) { asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// one staged query record at shared address a: the query words and its threshold
template <int W>
__device__ __forceinline__ QRec<W> lds_qrec(uint32_t a) {
  QRec<W> r;
  if constexpr (W == 1) {
    const uint4 v = lds_u4(a);
    r.qw[0] = v.x; r.qw[1] = v.y; r.tau = v.z;
  } else {
#pragma unroll
    for (int j = 0; j < W / 2; ++j) {
      const uint4 v = lds_u4(a + 16 * j);
      r.qw[4 * j] = v.x; r.qw[4 * j + 1] = v.y; r.qw[4 * j + 2] = v.z; r.qw[4 * j + 3] = v.w;
    }
    r.tau = lds_u32(a + 8 * W);
  }
  return r;
}

// One code (position j of table t, bucket order) of a queued hit against staged query qq of this warp: the exact distance
// against the staged record, and - for the few that really beat the threshold - de-duplication and append.
template <int W, int QS>
__device__ __noinline__ void bmih_check_code(const BmihParams* pp, uint32_t t, uint32_t j, uint32_t qq, CodeRegs<W> c) {
  BvWarp* w = &bv_warp[threadIdx.x >> 5];
  uint32_t* rec = w->qrec + qq * QS;
  const uint32_t tau = *(volatile uint32_t*)&rec[2 * W];
  const uint32_t d = hamming_exact<W>(c.w, rec);
  if (d > tau) return;
  const uint32_t qid = w->qid[qq];
#if VC_KEY_SUBST
  // d and tau are relative to the staged record: the substring of table t is not part of them (bmih_verify_kernel)
  const uint32_t r_sub = pp->scan_mode ? 0u : rec[2 * W + 1];
  bmih_append_impl<W>(pp, qid, t, d + r_sub, j, c, rec, tau + r_sub, pp->queries + (size_t)qid * 2 * W, r_sub);
#else
  bmih_append_impl<W>(pp, qid, t, d, j, c, rec, tau);
#endif
}
// The warp's queued hits, 32 at a time (called by all lanes).  A hit names a warp step of the item, a lane and a staged
// query: "one of the C codes that lane held in that step passed the filter for that query".  The draining lane reads those C
// codes again (they were streamed through this SM a moment ago: L2 hits, same addresses as bmih_verify_kernel's load_step)
// and checks each exactly - so the distance loop itself never has to find out which code it was.
template <int W, int U4, int QS>
__device__ __noinline__ void bmih_drain_hits(const BmihParams* pp, uint32_t t, uint32_t a0, uint32_t c0, uint32_t c1) {
  constexpr int C = 2 * U4 / W;
  BvWarp* w = &bv_warp[threadIdx.x >> 5];
  const uint32_t lane = threadIdx.x & 31;
  const uint4* src = reinterpret_cast<const uint4*>(bv_codes[t] + (size_t)a0 * W);
  const uint32_t u4_last = ((c1 - a0) * W + 1) / 2 - 1;
  __syncwarp();
  const uint32_t n = min(*(volatile uint32_t*)&w->hitn, (uint32_t)kBmihHitQ);
  for (uint32_t i = lane; i < n; i += 32) {
    const uint32_t e = w->hitq[i];
    const uint32_t step = e >> 10, ln = (e >> 5) & 31u, qq = e & 31u;
    const uint32_t base = a0 + step * (32 * C);
    const uint32_t u4_base = (base - a0) * W / 2;
    CodeRegs<W> code[C];
#pragma unroll
    for (int u = 0; u < U4; ++u) {
      uint4 v;
      if constexpr (W == 1) {
        v = __ldg(src + min(u4_base + u * 32 + ln, u4_last));
        code[2 * u].w[0] = v.x; code[2 * u].w[1] = v.y; code[2 * u + 1].w[0] = v.z; code[2 * u + 1].w[1] = v.w;
      } else if constexpr (W == 2) {
        v = __ldg(src + min(u4_base + u * 32 + ln, u4_last));
        code[u].w[0] = v.x; code[u].w[1] = v.y; code[u].w[2] = v.z; code[u].w[3] = v.w;
      } else {
        const int cc = u / 2, h = u % 2;
        v = __ldg(src + min(u4_base + 2 * (cc * 32 + ln), u4_last - 1) + h);
        code[cc].w[4 * h + 0] = v.x; code[cc].w[4 * h + 1] = v.y; code[cc].w[4 * h + 2] = v.z; code[cc].w[4 * h + 3] = v.w;
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const uint32_t j = base + (W == 1 ? 2 * ((c / 2) * 32 + ln) + (c & 1) : c * 32 + ln);
      if (j >= c0 && j < c1) bmih_check_code<W, QS>(pp, t, j, qq, code[c]);
    }
  }
  __syncwarp();
  if (lane == 0) w->hitn = 0;
  __syncwarp();
}

// Warp-granular: every warp of the persistent grid pulls its own work items (no block barriers), keeps the
// item's queries in its slice of shared memory and streams the item's codes 32 lanes x C codes at a time.
// The distance loop only FILTERS (one POPC per 64 bits when PREFILTER) the C codes a lane holds against one staged query; a
// lane whose minimum passes pushes one 16-bit word - (warp step, lane, staged query) - onto the warp's hit queue and moves on:
// no POPC, no call and no per-code work inside the divergent branch (the POPC pipe is what bounds this kernel, and a warp
// instruction costs it the same with one active lane as with 32).  The queue is worked off 32 hits at a time
// (bmih_drain_hits: exact re-check of those C codes, de-duplication, append - with full lanes), once a warp's worth has
// gathered and before the item is left; it holds what a whole warp step can add, so memory stays bounded whatever the data.
template <int W, bool PREFILTER, int U4>
__global__ void __launch_bounds__(kBmihThreads, U4 > 4 ? 2 : (W == 1 ? VC_VERIFY_CTAS : VC_VERIFY_CTAS_W2)) bmih_verify_kernel(const __grid_constant__ BmihParams p) {
  using Cfg = BmihCfg<W, U4>;
  constexpr int C = Cfg::C, QS = Cfg::QS;
  constexpr int NW = kBmihThreads / 32;
  constexpr uint32_t WSTEP = 32 * C;                       // codes per warp step
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  if (tid < p.m) { bv_codes[tid] = p.tables[tid].codes; bv_ids[tid] = p.tables[tid].ids; }
  if (tid < NW) { bv_pairs[tid] = 0; bv_warp[tid].hitn = 0; }
  __syncthreads();
  static_assert(QS <= kBmihQSMax, "staging area too small");
  const uint32_t wb = smem_u32(&bv_warp[tid >> 5]);       // this warp's state: every hot access below is [wb + constant (+ index)]
  // Codes not found before this step have substring distance >= r + 1 in the tables before t_begin and >= r in the
  // others: distance >= lb.  A query whose k-th best is already AT that distance can only improve by ties with a
  // smaller id - and a bucket is in ascending id order, so that query is done with a bucket at the first id >= its
  // k-th id (the last step of a search, table 0 of radius 3 at 1 B codes, scans ~30 % of every bucket instead of all).
  const uint32_t lb = p.m * p.r_lo + p.t_begin;
  const uint32_t n_items = *p.n_items;
  // Work items are claimed two ahead, all in registers: the descriptor of the next item (one word per lane 0..4) and the
  // claim of the one after it are in flight while the warp works on the current item, and are first touched at the top
  // of the next iteration.
  uint32_t claim = 0;
  if (lane == 0) claim = atomicAdd(p.item_cursor, 1u);
  uint32_t it_next = __shfl_sync(0xffffffffu, claim, 0);
  uint32_t ndw = 0;
  if (lane < 5 && it_next < n_items) ndw = reinterpret_cast<const uint32_t*>(p.items + it_next)[lane];
  if (lane == 0) claim = atomicAdd(p.item_cursor, 1u);
  for (;;) {
    const uint32_t it = it_next;
    if (it >= n_items) break;
    const uint32_t t = __shfl_sync(0xffffffffu, ndw, 0);
    const uint32_t c0 = __shfl_sync(0xffffffffu, ndw, 1), c1 = __shfl_sync(0xffffffffu, ndw, 2);
    const uint32_t qbeg = __shfl_sync(0xffffffffu, ndw, 3), qn = __shfl_sync(0xffffffffu, ndw, 4);
    it_next = __shfl_sync(0xffffffffu, claim, 0);
    ndw = 0;
    if (lane < 5 && it_next < n_items) ndw = reinterpret_cast<const uint32_t*>(p.items + it_next)[lane];
    if (lane == 0) claim = atomicAdd(p.item_cursor, 1u);
    const uint32_t a0 = W == 1 ? (c0 & ~1u) : c0;                      // 16-byte aligned start
    const uint4* src = reinterpret_cast<const uint4*>(bv_codes[t] + (size_t)a0 * W);
    const uint32_t u4_end = ((c1 - a0) * W + 1) / 2;                   // 16-byte units of the item
    const uint32_t u4_last = u4_end - 1;
    CodeRegs<W> code[C];
    auto load_step = [&](uint32_t base) {
      const uint32_t u4_base = (base - a0) * W / 2;
#pragma unroll
      for (int u = 0; u < U4; ++u) {
        // lanes past the end of the item re-read its last 16 bytes (no predicate, no zero fill): what they find is
        // dropped by the range test of the hit path
        uint4 v;
        if constexpr (W == 1) {
          const uint32_t idx = min(u4_base + u * 32 + lane, u4_last);
          v = ld_stream_u4(src + idx);
          code[2 * u].w[0] = v.x; code[2 * u].w[1] = v.y; code[2 * u + 1].w[0] = v.z; code[2 * u + 1].w[1] = v.w;
        } else if constexpr (W == 2) {
          const uint32_t idx = min(u4_base + u * 32 + lane, u4_last);
          v = ld_stream_u4(src + idx);
          code[u].w[0] = v.x; code[u].w[1] = v.y; code[u].w[2] = v.z; code[u].w[3] = v.w;
        } else {
          const int cc = u / 2, h = u % 2;
          const uint32_t idx = min(u4_base + 2 * (cc * 32 + lane), u4_last - 1) + h;
          v = ld_stream_u4(src + idx);
          code[cc].w[4 * h + 0] = v.x; code[cc].w[4 * h + 1] = v.y; code[cc].w[4 * h + 2] = v.z; code[cc].w[4 * h + 3] = v.w;
        }
      }
    };
    load_step(a0);                                     // the first codes travel while the queries are staged
    {
      // ... and so do the next kPfDist - 1 steps, into L2 (two lines per lane and step)
#pragma unroll
      for (int s1 = 1; s1 < kPfDist; ++s1) {
        const uint32_t u = (uint32_t)s1 * 128 + (lane & 15) * 8;
        if (lane < 16 && u < u4_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + u));
      }
    }
    __syncwarp();                                      // previous item's readers are done with the warp's slice
    {
      BvWarp* w = &bv_warp[tid >> 5];
      for (uint32_t e = lane; e < qn * QS; e += 32) {
        const uint32_t i = e / QS, x = e % QS;
        const uint32_t qid = p.qlist[qbeg + i];
        if (x == 0) w->qid[i] = qid;
        w->qrec[e] = x < 2 * W ? p.queries[(size_t)qid * 2 * W + x] : (x == 2 * W ? __ldcg(&p.gtau[qid]) : 0u);
      }
#if VC_KEY_SUBST
      // Every code of the item has the bucket's key as its substring t, so that substring adds the same r_own = popc(query
      // substring ^ key) to every distance.  With the key written into the staged query the XOR is zero there: the one-POPC
      // lower bound no longer loses those 16 bits to the OR with their partner substring, and is tested against tau - r_own (a
      // sharper filter at no cost per test: 128-bit codes at radius 4, tau = 37: a fifth instead of two thirds of the
      // warp-records contain a code that passes it).  The hit path adds r_own back and takes the query's own words for the
      // first-discoverer test; a threshold below r_own is clamped to 0, which can only let codes through that the append path
      // then handles with their true distance.
      static_assert(QS >= 2 * W + 2, "no spare word in the staged record");
      if (!p.scan_mode) {
        __syncwarp();
        if (lane < qn) {
          uint32_t* rec = w->qrec + lane * QS;
          const uint32_t key = substring<W>(reinterpret_cast<const uint32_t*>(bv_codes[t] + (size_t)c0 * W), t, p.sbits);
          const uint32_t off = t * p.sbits, wi = off >> 5, sh = off & 31;
          const uint32_t mask = (p.sbits == 32 ? 0xFFFFFFFFu : ((1u << p.sbits) - 1u)) << sh;
          const uint32_t qw = rec[wi];
          const uint32_t r_own = __popc((qw ^ (key << sh)) & mask);
          rec[wi] = (qw & ~mask) | (key << sh);
          const uint32_t tau0 = rec[2 * W];
          rec[2 * W] = tau0 >= r_own ? tau0 - r_own : 0u;
          rec[2 * W + 1] = r_own;
        }
      }
#endif
      uint32_t my_cut = 0xFFFFFFFFu;
      if (!p.scan_mode && lane < qn) {
        const uint64_t tk = __ldcg(&p.gtaukey[p.qlist[qbeg + lane]]);
        if ((uint32_t)(tk >> 32) == lb) my_cut = (uint32_t)tk;
      }
      w->cut[lane] = my_cut;
    }
    const bool use_cut = __any_sync(0xffffffffu, lds_u32(wb + kBvCut + 4 * lane) != 0xFFFFFFFFu);
    uint32_t qlive = qn;                               // staged queries still interested in the rest of the item
    uint32_t item_pairs = use_cut ? 0u : (c1 - c0) * qn;       // tests of this item (no query leaves early: counted once; < 2^32: cpi * 32)
    __syncwarp();
    for (uint32_t base = a0; base < c1; base += WSTEP) {
      if (use_cut) {
        // queries whose k-th id lies at or before the first id of this step are done with the bucket: the staged list
        // is compacted (each query leaves once), the distance loop below stays dense
        BvWarp* w = &bv_warp[tid >> 5];
        const uint32_t fid = __ldg(&bv_ids[t][max(base, c0)]);
        const bool keep = lane < qlive && w->cut[lane] > fid;
        const uint32_t alive = __ballot_sync(0xffffffffu, keep);
        if (!alive) break;
        const uint32_t live_mask = qlive >= 32 ? 0xFFFFFFFFu : ((1u << qlive) - 1u);
        if (alive != live_mask) {
          // The staged records are PARTITIONED, not dropped: the queries that go on move to the front, the ones that are done
          // with this bucket behind them (slots >= the new qlive are not tested any more, but hits queued for them earlier in
          // the item are still to be checked against their records), and the queued hits are renumbered accordingly.
          const bool mine = lane < qlive;
          const uint32_t gone = live_mask & ~alive;
          const uint32_t pos = keep ? __popc(alive & ((1u << lane) - 1u)) : __popc(alive) + __popc(gone & ((1u << lane) - 1u));
          uint32_t rec[QS];
#pragma unroll
          for (int i = 0; i < QS; ++i) rec[i] = mine ? w->qrec[lane * QS + i] : 0u;
          const uint32_t mq = mine ? w->qid[lane] : 0u, mc = mine ? w->cut[lane] : 0u;
          __syncwarp();
          if (mine) {
#pragma unroll
            for (int i = 0; i < QS; ++i) w->qrec[pos * QS + i] = rec[i];
            w->qid[pos] = mq; w->cut[pos] = mc;
          }
          w->perm[lane] = (uint8_t)(mine ? pos : lane);
          __syncwarp();
          if (VC_HIT_QUEUE) {
            const uint32_t nh = min(*(volatile uint32_t*)&w->hitn, (uint32_t)kBmihHitQ);
            for (uint32_t i = lane; i < nh; i += 32) { const uint32_t e = w->hitq[i]; w->hitq[i] = (uint16_t)((e & ~31u) | w->perm[e & 31u]); }
          }
          qlive = __popc(alive);
          __syncwarp();
        }
        item_pairs += (min(c1, base + WSTEP) - max(base, c0)) * qlive;
      }
      if constexpr (kPfDist > 0) {
        // the codes kPfDist steps ahead on their way into L2 (a warp step is 2 KB = 16 lines; no registers needed)
        const uint32_t ahead = (base - a0 + kPfDist * WSTEP) * W / 2;
        if (lane < 16 && ahead + lane * 8 < u4_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + ahead + lane * 8));
      }
      // keep the staged thresholds current: other warps (and, sharded, other GPUs) lower them all the time, and at
      // small radii - where the candidates are near neighbours by construction - a stale tau sends a large share
      // of the codes down the slow path.  The load is issued here and consumed after this step's math.
      uint32_t fresh_tau = kInfDist;
      const bool refresh = VC_TAU_EVERY == 1 || (((base - a0) / WSTEP) % VC_TAU_EVERY) == VC_TAU_EVERY - 1;
      if (refresh && lane < qlive) fresh_tau = __ldcg(&p.gtau[lds_u32(wb + kBvQid + 4 * lane)]);
      // one staged query (record at shared address qa) against this thread's C codes: the minimum of the (lower-bound)
      // distances decides; a lane whose minimum passes queues (step, lane, query slot) and moves on
      const uint32_t stepq = (((base - a0) / WSTEP) << 10) | (lane << 5);
      auto test_query = [&](const QRec<W>& cur, uint32_t q) {
        const uint32_t tau = cur.tau;
        uint32_t mn;
        if constexpr (C >= 3) {
          mn = PREFILTER ? hamming_lower_bound<W>(code[0].w, cur.qw) : hamming_exact<W>(code[0].w, cur.qw);
#pragma unroll
          for (int c = 1; c < C; ++c) mn = min(mn, PREFILTER ? hamming_lower_bound<W>(code[c].w, cur.qw) : hamming_exact<W>(code[c].w, cur.qw));
        } else {
          mn = 0xFFFFFFFFu;
#pragma unroll
          for (int c = 0; c < C; ++c) mn = min(mn, PREFILTER ? hamming_lower_bound<W>(code[c].w, cur.qw) : hamming_exact<W>(code[c].w, cur.qw));
        }
#if VC_HIT_QUEUE
        if (mn <= tau) {
          // the warp's queue is found from the thread index INSIDE the branch (the empty asm keeps the compiler from hoisting
          // the address out of the loop, where it would cost a register - or a reload from local memory - per record)
          uint32_t tw = threadIdx.x >> 5;
          asm volatile("" : "+r"(tw));
          BvWarp* w = &bv_warp[tw];
          w->hitq[atomicAdd(&w->hitn, 1u)] = (uint16_t)(stepq | q);
        }
#else
        if (mn <= tau) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            if ((PREFILTER ? hamming_lower_bound<W>(code[c].w, cur.qw) : hamming_exact<W>(code[c].w, cur.qw)) <= tau) {
              const uint32_t j = base + (W == 1 ? 2 * ((c / 2) * 32 + lane) + (c & 1) : c * 32 + lane);
              if (j >= c0 && j < c1) bmih_check_code<W, QS>(&p, t, j, q, code[c]);
            }
          }
        }
#endif
      };
      // two records in flight, ping-pong: the next record's LDS overlaps this record's math, without register moves
      {
        uint32_t qa = wb;
        QRec<W> ra = lds_qrec<W>(qa), rb;
#pragma unroll 1
        for (uint32_t q = 0; q < qlive; q += 2, qa += 2 * (4 * QS)) {
          if (q + 1 < qlive) rb = lds_qrec<W>(qa + 4 * QS);
          test_query(ra, q);
          if (q + 1 < qlive) {
            if (q + 2 < qlive) ra = lds_qrec<W>(qa + 2 * (4 * QS));
            test_query(rb, q + 1);
          }
        }
      }
      // the next step's codes start travelling now: the registers are free, and the rest of this step and the head of the next
      // one hide (part of) the latency
      if (base + WSTEP < c1) load_step(base + WSTEP);
      __syncwarp();
      if (refresh) {
        if (lane < qlive) {
#if VC_KEY_SUBST
          const uint32_t r_own = lds_u32(wb + lane * (4 * QS) + 4 * (2 * W + 1));
          const uint32_t ft = fresh_tau >= r_own ? fresh_tau - r_own : 0u;
#else
          const uint32_t ft = fresh_tau;
#endif
          if (ft < lds_u32(wb + lane * (4 * QS) + 8 * W)) atoms_min_u32(wb + lane * (4 * QS) + 8 * W, ft);
        }
        __syncwarp();
      }
      // hits of this step: worked off once a warp's worth has gathered, and before the item (its staged queries) is left
      if (VC_HIT_QUEUE) {
        const uint32_t hitn = lds_u32(wb + kBvHitN);
        if (hitn >= 32u || (hitn && base + WSTEP >= c1)) bmih_drain_hits<W, U4, QS>(&p, t, a0, c0, c1);
      }
    }
    if (VC_HIT_QUEUE && lds_u32(wb + kBvHitN)) bmih_drain_hits<W, U4, QS>(&p, t, a0, c0, c1);      // the id cut left the loop early
    if (lane == 0) bv_pairs[tid >> 5] += item_pairs;
  }
  if (lane == 0 && p.exec_pairs && bv_pairs[tid >> 5]) atomicAdd(p.exec_pairs, bv_pairs[tid >> 5]);
}

// ---- 4. settle: per query after a step --------------------------------------------------------------------
// sort the candidate buffer, keep k, refresh the exact local threshold and the distance histogram of what is
// kept; the histogram row is also copied to `xhist`, which the decide kernel reads - possibly after it has been
// summed over all id-shards (GPUs) by the caller's all-reduce.
template <int W>
__global__ void __launch_bounds__(256) bmih_settle_kernel(const BmihParams p, const uint32_t* list, uint32_t n_list, uint32_t* xhist, const XchgDev x) {
  __shared__ uint64_t buf[kBmihSort];
  __shared__ uint32_t cnt;
  __shared__ uint64_t s_tau;
  constexpr int HB = BmihCfg<W>::HB;
  constexpr uint32_t NB = 64 * W + 1;                     // distance bins 0 .. 64 W (HB leaves room for the two scratch words of topk_compact_block)
  __shared__ uint32_t sh[HB];
  const uint32_t tid = threadIdx.x;
  if (blockIdx.x >= n_list) return;                       // (never true with a peer exchange: the grid is exactly n_list CTAs)
  const uint32_t q = list[blockIdx.x];
  const uint32_t raw = p.gcnt[q];
  const uint32_t n = min(raw, p.cap);
  uint64_t* gb = p.gbuf + (size_t)q * p.cap;
  uint64_t tk;
  if (n <= (uint32_t)kBmihSort) {
    for (uint32_t i = tid; i < n; i += 256) buf[i] = gb[i];
    if (tid == 0) cnt = n;
    __syncthreads();
    tk = topk_compact_block(buf, &cnt, kBmihSort, p.k, sh, NB, tid, 256);
  } else {
    // more candidates than fit in shared memory (large k): folded in 256 at a time, the buffer compacted to its k best
    // whenever the next 256 might not fit; entries that no longer beat the running k-th key are dropped on the way in
    if (tid == 0) { cnt = 0; s_tau = kEmptyKey; }
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 256) {
      if (cnt + 256 > (uint32_t)kBmihSort) {                   // uniform: cnt is stable between barriers
        const uint64_t t2 = topk_compact_block(buf, &cnt, kBmihSort, p.k, sh, NB, tid, 256);
        if (tid == 0) s_tau = t2;
        __syncthreads();
      }
      if (base + tid < n) {
        const uint64_t key = gb[base + tid];
        if (key < s_tau) buf[atomicAdd(&cnt, 1u)] = key;
      }
      __syncthreads();
    }
    tk = topk_compact_block(buf, &cnt, kBmihSort, p.k, sh, NB, tid, 256);
  }
  const uint32_t kept = cnt;
  for (uint32_t i = tid; i < kept; i += 256) gb[i] = buf[i];
  // per-distance histogram of what is kept -> xhist (for the decide kernel / the cross-shard sum); its prefix
  // sums -> ghist (the cumulative counts the append path maintains)
  __syncthreads();
  for (uint32_t i = tid; i < HB; i += 256) sh[i] = 0;
  __syncthreads();
  for (uint32_t i = tid; i < kept; i += 256) atomicAdd(&sh[(uint32_t)(buf[i] >> 32)], 1u);
  __syncthreads();
  // the last bin is not a distance (distances end at 64 W): it carries "this query's buffer overflowed on this shard", so
  // that after the cross-shard sum every shard takes the query out of the batched search at the same step (bmih_decide_kernel)
  if (tid == 0) sh[HB - 1] = p.gflag[q] & 1u;
  __syncthreads();
  if (x.world) {
    // id-sharded over peer memory (xchg.cuh): the row goes straight into slot [rank] of every shard's window; xchg_sum_kernel
    // then leaves the sum over the shards in xhist
    const uint64_t off = x.slot_off + (uint64_t)x.rank * x.stride + (uint64_t)q * HB * 4;
    for (uint32_t i = tid; i < HB; i += 256)
      for (uint32_t g = 0; g < x.world; ++g) reinterpret_cast<uint32_t*>(x.peer[g] + off)[i] = sh[i];
  } else {
    for (uint32_t i = tid; i < HB; i += 256) xhist[(size_t)q * HB + i] = sh[i];
  }
  if (tid < 32) {                                         // one warp: lane l owns bins [l per, (l + 1) per)
    uint32_t* gc = p.ghist + (size_t)q * HB;
    constexpr uint32_t per = (NB + 31) / 32;
    uint32_t mine;
    uint32_t cum = warp_bins_excl(sh, NB, per, tid, mine);
    for (uint32_t j = 0; j < per; ++j) { const uint32_t d = tid * per + j; if (d < NB) { cum += sh[d]; gc[d] = cum; } }
  }
  if (tid == 0) {
    p.gcnt[q] = kept;
    p.gtaukey[q] = p.gglobkey ? min(tk, p.gglobkey[q]) : tk;
    if (tk != kEmptyKey) atomicMin(&p.gtau[q], (uint32_t)(tk >> 32));
  }
  if (x.world) xchg_publish(x);
}

// ---- 5. decide: stop rule and next active list ------------------------------------------------------------
// xhist[q][d] = number of known database codes at distance d from query q - this shard's kept candidates, or the
// sum over all shards.  tau = smallest d with at least k codes at distance <= d bounds the final k-th distance
// (of the whole database when summed), so it both tightens the filter and decides the stop: codes not found yet
// have substring distance >= r+1 in tables < t_end and >= r in the others, i.e. distance >= m*r + t_end, and the
// search can end as soon as that exceeds tau (search_worker.cc:204 made strict and m-aware; t_end == m gives
// the per-radius form d_k <= m*(r+1) - 1).  With summed histograms every shard takes the same decisions.
template <int W>
__global__ void bmih_decide_kernel(const BmihParams p, const uint32_t* list, uint32_t n_list, const uint32_t* xhist,
                                   uint32_t* any_overflow, uint32_t* n_likely, uint32_t* n_loose) {
  constexpr int HB = BmihCfg<W>::HB;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_list) return;
  const uint32_t q = list[i];
  uint32_t tau = kInfDist, cum = 0;
  for (uint32_t d = 0; d <= 64 * W; ++d) {
    cum += xhist[(size_t)q * HB + d];
    if (cum >= p.k) { tau = d; break; }
  }
  if (tau != kInfDist) atomicMin(&p.gtau[q], tau);
  const uint32_t r = p.radius;
  const bool level_done = p.t_end == p.m;                                                // a whole radius is finished
  bool stop = false;
  // Speculative thresholds (p.spec_tau, api.cu: the largest k-th distance of the previous batch on this index, plus one): the
  // filter started at min(bootstrap, spec_tau) instead of the bootstrap's looser bound.  If that guess was too small for this
  // query, fewer than k codes passed - and once every code within spec_tau has been enumerated (distance bound of what is left
  // > spec_tau) that is certain: the query leaves the batched search like an overflowed one and is answered by the per-query
  // kernel, exactly.  A query with k codes within spec_tau never notices: nothing at or below its k-th distance was filtered.
  const bool spec_failed = p.spec_tau != kInfDist && p.max_radius < 0 && tau == kInfDist && p.spec_tau + 1 <= p.m * r + p.t_end;
  const uint32_t tau_likely = min(tau, p.spec_tau);                                      // where the query will end if the guess holds
  if (level_done) {
    stop = r >= p.sbits;
    if (p.max_radius >= 0) stop = stop || r >= (uint32_t)p.max_radius;
  }
  if (p.max_radius < 0) stop = stop || (tau != kInfDist && tau + 1 <= p.m * r + p.t_end);
  if (spec_failed) { atomicOr(&p.gflag[q], 1u); *any_overflow = 1; stop = true; atomicAdd(p.spec_fail, 1u); }
  p.gradius[q] = r;
  for (uint32_t rr = p.r_lo; rr <= r; ++rr)
    p.gprobes[q] += (unsigned long long)(p.t_end - p.t_begin) * c_binom[p.sbits][rr];    // n_sub_reads_ of this step
  // candidate buffer overflowed (here or, id-sharded, on any shard: the flag bin is part of the summed histogram): the query
  // leaves the batched search now - on every shard at the same step - and is answered by the per-query kernel afterwards
  if (xhist[(size_t)q * HB + HB - 1]) { atomicOr(&p.gflag[q], 1u); *any_overflow = 1; stop = true; }
  // would this query stop somewhere inside the next radius even if tau did not improve any more?
  if (!stop && level_done && tau_likely != kInfDist && tau_likely + 1 <= p.m * (r + 1) + p.m) atomicAdd(n_likely, 1u);
  if (stop) atomicOr(&p.gflag[q], 2u);
  else {
    p.next_active[atomicAdd(p.n_next, 1u)] = q;
    // the next step probes at radius r (more tables of this radius) or r + 1: is this query's threshold, less that radius, still
    // too loose for the lower-bound filter?
    const uint32_t r_next = level_done ? r + 1 : r;
    if (__ldcg(&p.gtau[q]) >= p.pf_tau + r_next) atomicAdd(n_loose, 1u);
  }
}

// ---- id-sharded search: a bound on the k-th KEY of the whole database ------------------------------------------
// The all-reduced distance histograms give every shard the k-th DISTANCE tau of the whole database.  When the next step
// can only add codes at exactly that distance (lb_next == tau, see bmih_verify_kernel), what matters is the k-th ID among
// them - and a shard's own k-th key says little about it.  So the shards sum, per such query, the number of their kept
// codes below tau and a histogram of the ids (top 8 bits) of those at tau (same all-reduce callback); the first bin
// edge at which k codes are reached is a key that k codes of the database are below: candidates at or above it cannot be
// in the answer, whichever shard finds them.  With ids interleaved over the shards every shard then leaves its buckets
// at ~ the same fraction as a single GPU would.
constexpr int kIdBins = 256;
__global__ void bmih_idhist_kernel(const BmihParams p, const uint32_t* list, uint32_t n_list, uint32_t lb_next, uint32_t* idh) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_list) return;
  const uint32_t q = list[w];
  const uint32_t tau = p.gtau[q];
  if (tau != lb_next) return;
  uint32_t* row = idh + (size_t)q * (kIdBins + 1);
  const uint32_t n = min(p.gcnt[q], p.k);
  for (uint32_t i = lane; i < n; i += 32) {
    const uint64_t key = p.gbuf[(size_t)q * p.cap + i];
    const uint32_t d = (uint32_t)(key >> 32);
    if (d < tau) atomicAdd(&row[0], 1u);
    else if (d == tau) atomicAdd(&row[1 + ((uint32_t)key >> 24)], 1u);
  }
}
__global__ void bmih_idcut_kernel(const BmihParams p, const uint32_t* list, uint32_t n_list, uint32_t lb_next, const uint32_t* idh) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_list) return;
  const uint32_t q = list[i];
  if (p.gtau[q] != lb_next) return;
  const uint32_t* row = idh + (size_t)q * (kIdBins + 1);
  uint32_t cum = row[0];
  uint64_t bound = kEmptyKey;
  for (uint32_t b = 0; b < (uint32_t)kIdBins; ++b) {
    cum += row[1 + b];
    if (cum >= p.k) { bound = b + 1 == (uint32_t)kIdBins ? pack_key(lb_next + 1, 0) : pack_key(lb_next, (b + 1) << 24); break; }
  }
  if (bound < p.gglobkey[q]) p.gglobkey[q] = bound;
  if (bound < p.gtaukey[q]) p.gtaukey[q] = bound;
}

// largest final k-th distance of the queries the batched search answered itself (the next batch's speculative threshold)
__global__ void bmih_maxtau_kernel(const BmihParams p, uint32_t* out) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v = 0;
  if (q < p.nq && !(p.gflag[q] & 1u)) v = p.gtau[q];                 // kInfDist (no k codes found at all) makes the guess void
  v = __reduce_max_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && v) atomicMax(out, v);
}

// per-query state at the start of a search
__global__ void bmih_init_kernel(const BmihParams p, uint32_t* active0) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.nq) return;
  p.gcnt[q] = 0; p.gtau[q] = p.spec_tau; p.gtaukey[q] = kEmptyKey; if (p.gglobkey) p.gglobkey[q] = kEmptyKey; p.gflag[q] = 0; p.gradius[q] = 0;
  p.gprobes[q] = 0; p.gcands[q] = 0;
  active0[q] = q;
}

// Bootstrap of the distance thresholds: before level 0 nothing is known about the k-th distance, and with
// tau = infinity every member of the query's own buckets (N / 2^s codes per table) would have to be buffered.
// One warp per query looks at up to kBmihSample codes of those buckets (table 0 first), histograms their
// distances (each code once: the first-discoverer rule at radius 0) and sets tau[q] to the k-th smallest -
// a valid bound, since these are real database codes that level 0 will find again.
constexpr uint32_t kBmihSample = 16384;    // at least; 16 * k when that is more (16 K instead of 4 K: -0.4 ms in the first step)
template <int W>
__global__ void __launch_bounds__(256) bmih_bootstrap_kernel(const BmihParams p, uint32_t* xh /* id-sharded: the sample histograms, to be summed over the shards */) {
  constexpr int HB = BmihCfg<W>::HB;
  __shared__ uint32_t s_hist[8][HB];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t q = blockIdx.x * 8 + warp;
  if (q >= p.nq) return;
  uint32_t* h = s_hist[warp];
  for (uint32_t i = lane; i < HB; i += 32) h[i] = 0;
  __syncwarp();
  uint32_t qw[2 * W];
#pragma unroll
  for (int i = 0; i < 2 * W; ++i) qw[i] = p.queries[(size_t)q * 2 * W + i];
  uint32_t taken = 0;
  const uint32_t sample = p.boot_sample ? p.boot_sample : max(kBmihSample, 16u * p.k);
  for (uint32_t t = 0; t < p.m && taken < sample; ++t) {
    const uint32_t key = substring<W>(qw, t, p.sbits);
    const uint32_t start = p.tables[t].row_ptr[key], len = p.tables[t].row_ptr[key + 1] - start;
    const uint32_t take = min(len, sample - taken);
    const uint64_t* codes = p.tables[t].codes;
    // U codes per lane in flight: one load at a time made this kernel pure DRAM latency (512 round trips, 0.27 ms)
    constexpr int U = 8 / W;
    for (uint32_t j0 = 0; j0 < take; j0 += 32 * U) {
      uint64_t c[U][W];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t j = min(j0 + u * 32 + lane, take - 1);
#pragma unroll
        for (int i = 0; i < W; ++i) c[u][i] = __ldg(&codes[(size_t)(start + j) * W + i]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j0 + u * 32 + lane >= take) continue;
        uint32_t x[2 * W];
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < W; ++i) {
          x[2 * i] = (uint32_t)c[u][i] ^ qw[2 * i]; x[2 * i + 1] = (uint32_t)(c[u][i] >> 32) ^ qw[2 * i + 1];
          d += __popc(x[2 * i]) + __popc(x[2 * i + 1]);
        }
        bool first = true;                        // already counted from a lower table whose substring also matches?
        for (uint32_t t2 = 0; t2 < t; ++t2) first = first && substring<W>(x, t2, p.sbits) != 0;
        if (first) atomicAdd(&h[d], 1u);
      }
    }
    taken += take;
  }
  __syncwarp();
  if (xh) for (uint32_t i = lane; i < HB; i += 32) xh[(size_t)q * HB + i] = h[i];
  if (lane == 0) {
    uint32_t cum = 0;
    for (uint32_t d = 0; d <= 64 * W; ++d) {
      cum += h[d];
      if (cum >= p.k) { p.gtau[q] = min(d, p.spec_tau); break; }
    }
  }
}
// id-sharded search: the shards' samples are disjoint sets of real codes, so the k-th smallest distance of their union
// bounds the k-th distance of the whole database - G times the sample, a much tighter start than a shard's own
template <int W>
__global__ void bmih_boot_tau_kernel(const BmihParams p, const uint32_t* xh) {
  constexpr int HB = BmihCfg<W>::HB;
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.nq) return;
  uint32_t cum = 0;
  for (uint32_t d = 0; d <= 64 * W; ++d) {
    cum += xh[(size_t)q * HB + d];
    if (cum >= p.k) { atomicMin(&p.gtau[q], d); break; }
  }
}

// ---- brute-force scan through the verify kernel ---------------------------------------------------------------
// work items: code chunk c x query chunk qc, query chunk fastest (neighbouring items share their codes in L2)
__global__ void scan_items_kernel(BmihItem* items, uint64_t n, uint32_t cpi, uint32_t nq, uint32_t nqc, uint32_t n_items) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const uint32_t c = i / nqc, qc = i % nqc;
    BmihItem it;
    it.t = 0;
    it.c0 = c * cpi;
    it.c1 = (uint32_t)min(n, (uint64_t)it.c0 + cpi);
    const uint32_t qlo = (uint32_t)(((uint64_t)nq * qc) / nqc), qhi = (uint32_t)(((uint64_t)nq * (qc + 1)) / nqc);
    it.qbeg = qlo; it.qn = qhi - qlo;
    items[i] = it;
  }
}
// threshold bootstrap for the scan: k-th smallest distance among the first `sample` codes of the shard, one CTA per
// query.  All warps of the persistent grid start on the same few query chunks, so the first thresholds must already
// be tight enough for (codes in flight) x (pass rate) to stay far below the candidate buffer: a pilot pass over the
// first kBmihSample codes bounds the distance, the second pass histograms only what lies below that bound.
constexpr uint32_t kScanSample = 1u << 18;
template <int W>
__global__ void __launch_bounds__(256) scan_bootstrap_kernel(const BmihParams p, const uint64_t* codes, uint64_t n) {
  constexpr int HB = BmihCfg<W>::HB;
  __shared__ uint32_t s_hist[HB];
  __shared__ uint32_t s_bound;
  const uint32_t tid = threadIdx.x, q = blockIdx.x;
  uint32_t qw[2 * W];
#pragma unroll
  for (int i = 0; i < 2 * W; ++i) qw[i] = p.queries[(size_t)q * 2 * W + i];
  const uint32_t sample = (uint32_t)min(n, (uint64_t)max(kScanSample, 64u * p.k));
  const uint32_t pilot = min(sample, max(kBmihSample, 16u * p.k));
  uint32_t bound = 64 * W;
  for (int pass = 0; pass < 2; ++pass) {
    for (uint32_t i = tid; i < HB; i += 256) s_hist[i] = 0;
    __syncthreads();
    const uint32_t j1 = pass ? sample : pilot;      // the second pass counts the pilot's codes again: >= k below the bound
    for (uint32_t j = tid; j < j1; j += 256) {
      uint32_t d = 0;
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const uint64_t c = codes[(size_t)j * W + i];
        d += __popc((uint32_t)c ^ qw[2 * i]) + __popc((uint32_t)(c >> 32) ^ qw[2 * i + 1]);
      }
      if (d <= bound) atomicAdd(&s_hist[d], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t cum = 0, b = kInfDist;
      for (uint32_t d = 0; d <= 64 * W; ++d) {
        cum += s_hist[d];
        if (cum >= p.k) { b = d; break; }
      }
      s_bound = b;
    }
    __syncthreads();
    if (pass == 0) {
      if (s_bound == kInfDist || pilot == sample) break;
      bound = s_bound;
    }
    __syncthreads();
  }
  if (tid == 0 && s_bound != kInfDist) p.gtau[q] = s_bound;
}

// With a peer exchange (x.world > 0, id-sharded search) the row also goes straight into slot [rank] of every shard's window:
// the all-gather of the local top-k lists is this kernel's own stores (xchg.cuh), the merge kernel follows after the wait.
__global__ void bmih_finish_kernel(const BmihParams p, uint64_t* out_keys, vc_query_stats* stats, const XchgDev x) {
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  const uint32_t kept = min(p.gcnt[q], p.k);
  for (uint32_t i = tid; i < p.k; i += blockDim.x) {
    const uint64_t key = i < kept ? p.gbuf[(size_t)q * p.cap + i] : kEmptyKey;
    out_keys[(size_t)q * p.k + i] = key;
    if (x.world) {
      const uint64_t off = x.slot_off + (uint64_t)x.rank * x.stride + ((uint64_t)q * p.k + i) * 8;
      for (uint32_t g = 0; g < x.world; ++g) *reinterpret_cast<uint64_t*>(x.peer[g] + off) = key;
    }
  }
  if (stats && tid == 0) {
    vc_query_stats st;
    st.radius = p.gradius[q]; st.n_results = kept; st.probes = p.gprobes[q]; st.occupancy_tests = 0;
    st.candidates = p.gcands[q]; st.unique = 0;
    stats[q] = st;
  }
  if (x.world) xchg_publish(x);
}

// queries whose candidate buffer overflowed: listed for the per-query kernel, which writes their rows of the output itself
// ---- approximate mode on the batched path (api.cu: mih_approx_batched) ---------------------------------------
// search_K_approximate_nearest_neighbors (search_worker.cc:93-157) widens the radius until k * 20 DISTINCT candidates are
// known (knn_found_, :117-124) and returns the k best of them.  With long buckets that is radius 0 for nearly every query:
// the answer is then the fixed-radius-0 search, which the batched path does, and what remains is to know, per query,
// whether radius 0 really was enough and - for the statistics - how many distinct candidates it saw.  One CTA per query:
// the query's m own buckets; a bucket holds distinct codes, so the longest one is a lower bound and the sum an upper bound
// of the count; only when those do not decide (or the exact count is asked for) are the buckets streamed and the codes
// counted that no EARLIER table also finds at radius 0 (the first-discoverer rule).
// flag[q] = 1: radius 0 is not enough, the per-query kernel answers this query.
template <int W>
__global__ void __launch_bounds__(256) approx_radius0_kernel(const uint32_t* __restrict__ queries, const TableDev* __restrict__ tables,
                                                             uint32_t m, uint32_t sbits, uint32_t nq, unsigned long long need,
                                                             int count_always, unsigned long long* unique0, uint32_t* flag) {
  __shared__ uint32_t s_start[kMaxTables], s_len[kMaxTables];
  __shared__ uint32_t s_q[2 * W];
  __shared__ unsigned long long s_sum;
  __shared__ uint32_t s_max, s_count;
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  if (q >= nq) return;
  if (tid < 2 * W) s_q[tid] = queries[(size_t)q * 2 * W + tid];
  if (tid == 0) { s_sum = 0; s_max = 0; s_count = 0; }
  __syncthreads();
  if (tid < m) {
    uint32_t st = 0, ln = 0;
    table_lookup(tables[tid], substring<W>(s_q, tid, sbits), st, ln);
    s_start[tid] = st; s_len[tid] = ln;
    atomicAdd(&s_sum, (unsigned long long)ln);
    atomicMax(&s_max, ln);
  }
  __syncthreads();
  if (!count_always) {
    if ((unsigned long long)s_max >= need) { if (tid == 0) { unique0[q] = ~0ull; flag[q] = 0; } return; }
    if (s_sum < need) { if (tid == 0) { unique0[q] = s_sum; flag[q] = 1; } return; }
  }
  uint32_t qw[2 * W];
#pragma unroll
  for (int i = 0; i < 2 * W; ++i) qw[i] = s_q[i];
  uint32_t mine = 0;
  for (uint32_t t = 0; t < m; ++t) {
    const uint64_t* bc = tables[t].codes + (size_t)s_start[t] * W;
    const uint32_t len = s_len[t];
    if (t == 0) { if (tid == 0) mine += len; continue; }      // nothing precedes table 0
    for (uint32_t j = tid; j < len; j += 256) {
      uint32_t x[2 * W];
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const uint64_t c = __ldcs(bc + (size_t)j * W + w);
        x[2 * w] = (uint32_t)c ^ qw[2 * w]; x[2 * w + 1] = (uint32_t)(c >> 32) ^ qw[2 * w + 1];
      }
      bool first = true;
      for (uint32_t t2 = 0; t2 < t; ++t2) first = first && substring<W>(x, t2, sbits) != 0;
      mine += first ? 1u : 0u;
    }
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((tid & 31) == 0 && mine) atomicAdd(&s_count, mine);
  __syncthreads();
  if (tid == 0) { unique0[q] = s_count; flag[q] = (unsigned long long)s_count < need ? 1u : 0u; }
}
// statistics of the queries that ended at radius 0: the distinct-candidate count (the batched search wrote the rest)
__global__ void approx_stats_kernel(vc_query_stats* stats, const unsigned long long* unique0, const uint32_t* flag, uint32_t nq) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq && !flag[q]) stats[q].unique = unique0[q];
}

__global__ void bmih_redo_list_kernel(const uint32_t* gflag, uint32_t nq, uint32_t* list, uint32_t* n_list) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq && (gflag[q] & 1u)) list[atomicAdd(n_list, 1u)] = q;
}

}  // namespace vc
