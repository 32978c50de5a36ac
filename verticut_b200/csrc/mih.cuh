// MIH k-NN query kernel (K2-K5): one CTA per query.
//
// Replaces SearchWorker::find and its helpers, src/search_worker.cc:65-264:
//   radius loop r = 0,1,...                                   search_K_nearest_neighbors :159-218
//     every table t enumerates all bucket indices at Hamming distance exactly r from the
//     query's substring t                                     enumerate_entry :230-264
//     every member of a probed bucket is verified with XOR + POPC over the full code  :249-257
//     candidates are de-duplicated across tables                knn_found_ (std::map) :183-190
//     the k best are kept                                       std::priority_queue :192-197
//     stop test                                                 :201-207
// What changes on the GPU:
//   * enumeration: probe p of radius r is UNRANKED (combinatorial number system over the s bit
//     positions) so that each lane of a warp owns one probe; the warp then streams the 32 buckets
//     cooperatively with 128-bit loads (bucket members are contiguous in HBM, see build.cuh).
//   * de-duplication without memory traffic: a code reaches the result stage exactly once - from
//     the first table (lowest t) among those whose substring is closest to the query - which is
//     decided from the code itself (per-substring POPC), only for the few candidates that pass the
//     distance threshold.
//   * top-k: canonical (dist,id) order; per-warp staging buffers are folded into one sorted
//     per-query buffer under a shared-memory lock, so memory stays bounded whatever the ties.
//   * stop rule: m-aware and strict, d_k <= m*(r+1) - 1, which makes the result identical to a
//     full scan (SURVEY.md findings 6, 7).  This per-query kernel applies it to its own id-shard
//     (no exchange); the batched path (bmih.cuh) sums the shards' distance histograms once per step
//     and stops on the k-th distance of the whole database.
#pragma once
#include "build.cuh"
#include "../../include/verticut_gpu.h"

namespace vc {

constexpr int kMihThreads = 256;
constexpr int kMihWarps = kMihThreads / 32;
constexpr int kMihWbuf = 128;         // per-warp staging entries
constexpr int kMaxTables = 32;

__constant__ uint32_t c_binom[33][33];   // C(n, k), n, k <= 32 (fits 32 bits)

struct MihParams {
  const uint32_t* queries;   // [nq][2W]
  uint32_t nq, k;
  uint32_t m, sbits;
  uint32_t BUFM;             // per-query buffer entries: power of two >= k + kMihWbuf
  int approximate;           // stop when >= 20k distinct candidates were seen (search_worker.cc:136-137)
  int max_radius;            // >= 0: fixed radius; < 0: stop rule
  int table_steps;           // exact mode: 1 = test the stop rule after every table of a radius (d_k <= m*r + t), 0 = per radius
  const TableDev* tables;    // [m] in device memory
  const uint32_t* qsel;      // null: CTA b answers query b; else CTA b answers query qsel[b] (rows of queries / out_keys / stats are indexed by query)
  uint64_t* out_keys;        // [nq][k]
  vc_query_stats* stats;     // [nq] or null
};

// mask with r bits set among s positions, rank p in [0, C(s,r)) (combinatorial number system)
__device__ __forceinline__ uint32_t unrank_mask(uint32_t s, uint32_t r, uint32_t p) {
  uint32_t mask = 0;
  int c = (int)s - 1;
  for (uint32_t i = r; i > 0; --i) {
    while (c_binom[c][i] > p) --c;      // largest c with C(c, i) <= p
    mask |= 1u << c;
    p -= c_binom[c][i];
    --c;
  }
  return mask;
}

// substring t (sbits wide) of a W-word value held as 2W 32-bit words
template <int W>
__device__ __forceinline__ uint32_t substring(const uint32_t* x, uint32_t t, uint32_t sbits) {
  const uint32_t off = t * sbits;
  const uint32_t v = x[off >> 5] >> (off & 31);
  return sbits == 32 ? v : (v & ((1u << sbits) - 1));
}

// Folds a warp's staging buffer into the per-query buffer.  Called by all 32 lanes.
__device__ __forceinline__ void mih_flush(uint64_t* mbuf, volatile uint64_t* tau_key, volatile uint32_t* tau_dist,
                                          uint32_t* mcnt, uint32_t* lock, uint32_t BUFM, uint32_t k,
                                          const uint64_t* wbuf, uint32_t& wcnt, uint32_t lane) {
  if (wcnt == 0) return;
  if (lane == 0) { while (atomicCAS(lock, 0u, 1u) != 0u) __nanosleep(32); }
  __syncwarp();
  __threadfence_block();
  uint32_t c = *(volatile uint32_t*)mcnt;
  const uint64_t tau = *tau_key;
  for (uint32_t i = 0; i < wcnt; i += 32) {
    const bool pass = i + lane < wcnt && wbuf[i + lane] < tau;
    const uint32_t mask = __ballot_sync(0xffffffffu, pass);
    if (pass) mbuf[c + __popc(mask & ((1u << lane) - 1))] = wbuf[i + lane];
    c += __popc(mask);
  }
  __syncwarp();
  if (lane == 0) *mcnt = c;
  __syncwarp();
  const uint64_t ntau = topk_compact(mbuf, mcnt, BUFM, k, lane, 32, WarpSync());
  if (lane == 0) {
    *tau_key = ntau;
    if (ntau != kEmptyKey) atomicMin((uint32_t*)tau_dist, (uint32_t)(ntau >> 32));
    __threadfence_block();
    atomicExch(lock, 0u);
  }
  __syncwarp();
  wcnt = 0;
}

template <int W, bool APPROX>
__global__ void __launch_bounds__(kMihThreads, 3) mih_search_kernel(const MihParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* mbuf = (uint64_t*)smem_raw;
  uint64_t* wbuf_all = mbuf + p.BUFM;
  __shared__ uint64_t s_tau_key;
  __shared__ uint32_t s_tau_dist, s_mcnt, s_lock, s_stop;
  __shared__ uint32_t s_q[2 * W];
  __shared__ uint32_t s_qkey[kMaxTables];
  __shared__ unsigned long long s_probes, s_occ, s_cands, s_unique;
  __shared__ TableDev s_tab[kMaxTables];
  __shared__ uint32_t s_hist[64 * W + 32];      // distances of the distinct candidates that passed the threshold

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t q = p.qsel ? p.qsel[blockIdx.x] : blockIdx.x;
  const uint32_t m = p.m, sbits = p.sbits, k = p.k;

  if (tid < 2 * W) s_q[tid] = p.queries[(size_t)q * 2 * W + tid];
  if (tid < m) s_tab[tid] = p.tables[tid];
  for (uint32_t i = tid; i < 64 * W + 32; i += kMihThreads) s_hist[i] = 0;
  if (tid == 0) {
    s_tau_key = kEmptyKey; s_tau_dist = kInfDist; s_mcnt = 0; s_lock = 0; s_stop = 0;
    s_probes = 0; s_occ = 0; s_cands = 0; s_unique = 0;
  }
  __syncthreads();
  if (tid < m) s_qkey[tid] = substring<W>(s_q, tid, sbits);      // search_worker.cc:165-167
  __syncthreads();

  uint32_t qw[2 * W];
#pragma unroll
  for (int i = 0; i < 2 * W; ++i) qw[i] = s_q[i];
  uint64_t* wbuf = wbuf_all + (size_t)warp * kMihWbuf;
  uint32_t wcnt = 0;
  volatile uint32_t* v_tau_dist = &s_tau_dist;
  volatile uint64_t* v_tau_key = &s_tau_key;

  uint32_t radius = 0;
  unsigned long long my_probes = 0, my_occ = 0, my_cands = 0, my_unique = 0;

  const bool tsteps = p.table_steps != 0 && !APPROX && p.max_radius < 0;
  bool done = false;
  for (;; ++radius) {                                                // search_worker.cc:170
    const uint32_t per_table = c_binom[sbits][radius];              // C(s, r) probes per table
   for (uint32_t t0 = 0; t0 < m && !done;) {
    const uint32_t t1 = tsteps ? t0 + 1 : m;                         // tables [t0, t1) of this radius in one go
    const uint64_t total = (uint64_t)per_table * (t1 - t0);
    // few probes (radius 0, 1): every warp visits every bucket and streams its own eighth of it; otherwise
    // the warps take different groups of 32 probes
    const bool split = total < (uint64_t)kMihWarps * 32 * 2;
    for (uint64_t g = split ? 0 : (uint64_t)warp * 32; g < total; g += split ? 32 : (uint64_t)kMihWarps * 32) {
      // ---- one probe per lane ------------------------------------------------------------
      const uint64_t item = g + lane;
      uint32_t bt = 0, bstart = 0, blen = 0;
      if (item < total) {
        bt = t0 + (uint32_t)(item / per_table);
        const uint32_t pidx = (uint32_t)(item % per_table);
        const uint32_t key = s_qkey[bt] ^ unrank_mask(sbits, radius, pidx);      // :260 curr ^ (1 << len)
        table_lookup(s_tab[bt], key, bstart, blen);                             // :246 proxy get
        if (!split || warp == 0) {                                               // count each probe once
          if (s_tab[bt].sparse) { ++my_occ; my_probes += blen ? 1 : 0; }         // :238-245
          else ++my_probes;
          my_cands += blen;
        }
      }
      // ---- the warp streams the non-empty buckets one after the other -----------------------
      uint32_t todo = __ballot_sync(0xffffffffu, blen != 0);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t t = __shfl_sync(0xffffffffu, bt, src);
        const uint32_t start = __shfl_sync(0xffffffffu, bstart, src);
        const uint32_t len = __shfl_sync(0xffffffffu, blen, src);
        const uint64_t* bc = s_tab[t].codes;
        const uint32_t* bi = s_tab[t].ids;
        // 16-byte units: W == 1 -> two codes per unit (unit = codes 2u, 2u+1), else W/2 units per code.
        // Each lane keeps U independent 128-bit loads in flight per iteration.
        constexpr int CPL = W == 1 ? 2 : 1;                 // codes per 128-bit group
        constexpr int UPC = W == 1 ? 1 : W / 2;             // 128-bit loads per group
        constexpr int U = W == 4 ? 2 : 4;                   // groups per lane per iteration
        const uint32_t first = W == 1 ? (start & ~1u) : start;
        const uint32_t last = start + len;                  // exclusive
        const uint32_t jstep = 32 * CPL * U;
        for (uint32_t j0 = first + (split ? warp * jstep : 0); j0 < last; j0 += split ? kMihWarps * jstep : jstep) {
          uint32_t cw[U][CPL][2 * W];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint32_t j = j0 + (u * 32 + lane) * CPL;
            if (j < last) {
              const uint4* src4 = reinterpret_cast<const uint4*>(bc + (size_t)j * W);
              if constexpr (W == 1) {
                const uint4 v = ld_stream_u4(src4);
                cw[u][0][0] = v.x; cw[u][0][1] = v.y; cw[u][1][0] = v.z; cw[u][1][1] = v.w;
              } else {
#pragma unroll
                for (int h = 0; h < UPC; ++h) {
                  const uint4 v = ld_stream_u4(src4 + h);
                  cw[u][0][4 * h] = v.x; cw[u][0][4 * h + 1] = v.y; cw[u][0][4 * h + 2] = v.z; cw[u][0][4 * h + 3] = v.w;
                }
              }
            }
          }
          const uint32_t tau = *v_tau_dist;
          uint32_t dd[U][CPL];
          uint32_t mn = kInfDist;
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
              const uint32_t jj = j0 + (u * 32 + lane) * CPL + c;
              const bool in = jj >= start && jj < last;
              dd[u][c] = in ? hamming_exact<W>(cw[u][c], qw) : kInfDist;          // :253 compute_hamming_dist
              mn = min(mn, dd[u][c]);
            }
          // common case: nothing in this stretch beats the current k-th distance
          if (!__any_sync(0xffffffffu, APPROX ? mn != kInfDist : mn <= tau)) continue;
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
              const uint32_t jj = j0 + (u * 32 + lane) * CPL + c;
              const uint32_t d = dd[u][c];
              const bool in = d != kInfDist;
              bool pass = in && d <= tau;
              if (APPROX ? in : pass) {
                // first-discoverer test: emitted only by the lowest table among those whose substring
                // distance equals the minimum (which is `radius` for table t at this step)
                bool first_seen = true;
                uint32_t x[2 * W];
#pragma unroll
                for (int i = 0; i < 2 * W; ++i) x[i] = cw[u][c][i] ^ qw[i];
                for (uint32_t t2 = 0; t2 < m; ++t2) {
                  if (t2 == t) continue;
                  const uint32_t sd = __popc(substring<W>(x, t2, sbits));
                  if (sd < radius || (sd == radius && t2 < t)) { first_seen = false; break; }
                }
                if (APPROX && first_seen) ++my_unique;
                pass = pass && first_seen;
              }
              uint64_t key = 0;
              if (pass) {
                key = pack_key(d, bi[jj]);
                pass = key < *v_tau_key;
              }
              if (pass) {
                // exact running threshold: count the distance, lower tau once k codes are closer
                atomicAdd(&s_hist[d], 1u);
                const uint32_t tnow = *v_tau_dist;
                if (d < tnow) {
                  uint32_t cum = 0;
                  const uint32_t lim = min(tnow, (uint32_t)(64 * W + 1));
                  for (uint32_t b2 = 0; b2 < lim; ++b2) {
                    cum += *(volatile uint32_t*)&s_hist[b2];
                    if (cum >= k) { atomicMin((uint32_t*)&s_tau_dist, b2); break; }
                  }
                }
              }
              const uint32_t mask = __ballot_sync(0xffffffffu, pass);
              if (mask) {
                if (wcnt + 32 > kMihWbuf)
                  mih_flush(mbuf, v_tau_key, v_tau_dist, &s_mcnt, &s_lock, p.BUFM, k, wbuf, wcnt, lane);
                if (pass) wbuf[wcnt + __popc(mask & ((1u << lane) - 1))] = key;
                wcnt += __popc(mask);
                __syncwarp();
              }
            }
        }
      }
    }
    // ---- end of the step: fold every warp's staging buffer, then decide ---------------------------------
    mih_flush(mbuf, v_tau_key, v_tau_dist, &s_mcnt, &s_lock, p.BUFM, k, wbuf, wcnt, lane);
    if (APPROX) {
      for (int o = 16; o > 0; o >>= 1) my_unique += __shfl_xor_sync(0xffffffffu, my_unique, o);
      if (lane == 0 && my_unique) atomicAdd(&s_unique, my_unique);
      my_unique = 0;
    }
    __syncthreads();
    if (tid == 0) {
      bool stop = false;
      if (t1 == m) {                                                              // a whole radius is finished
        stop = radius >= sbits;                                                   // :170 radius <= s
        if (p.max_radius >= 0) stop = stop || radius >= (uint32_t)p.max_radius;
        else if (APPROX) stop = stop || s_unique >= (unsigned long long)k * VC_APPROXIMATE_FACTOR;   // :136-137
      }
      // strict, m-aware stop rule (:204): codes not found yet have substring distance >= r+1 in tables < t1 and
      // >= r in the others, i.e. distance >= m*r + t1 > d_k.  For t1 == m this is d_k <= m*(r+1) - 1.
      if (!APPROX && p.max_radius < 0) stop = stop || (s_mcnt == k && (uint32_t)(s_tau_key >> 32) + 1 <= m * radius + t1);
      s_stop = stop ? 1u : 0u;
    }
    __syncthreads();
    if (s_stop) done = true;
    t0 = t1;
   }
    if (done) break;
  }

  // ---- results and statistics ---------------------------------------------------------------------
  const uint32_t kept = s_mcnt;
  for (uint32_t i = tid; i < k; i += kMihThreads) p.out_keys[(size_t)q * k + i] = i < kept ? mbuf[i] : kEmptyKey;
  if (p.stats) {
    for (int o = 16; o > 0; o >>= 1) {
      my_probes += __shfl_xor_sync(0xffffffffu, my_probes, o);
      my_occ += __shfl_xor_sync(0xffffffffu, my_occ, o);
      my_cands += __shfl_xor_sync(0xffffffffu, my_cands, o);
    }
    if (lane == 0) { atomicAdd(&s_probes, my_probes); atomicAdd(&s_occ, my_occ); atomicAdd(&s_cands, my_cands); }
    __syncthreads();
    if (tid == 0) {
      vc_query_stats st;
      st.radius = radius; st.n_results = kept; st.probes = s_probes; st.occupancy_tests = s_occ;
      st.candidates = s_cands; st.unique = APPROX ? s_unique : 0;
      p.stats[q] = st;
    }
  }
}

}  // namespace vc
