// Linear scan (K6) + top-k selection (K5) + partial-list merge (K7).
//
// Replaces search_K_nearest_neighbors(int k) of the reference, src/linear_search.cc:39-64
// (for every code: compute_hamming_dist, max-heap of k) for a BATCH of queries, with the
// canonical tie rule (k smallest packed words dist<<32|id, ascending).
//
// Decomposition: the database shard is cut into `n_slices` contiguous slices, the query batch into
// `n_qtiles` tiles of QT queries.  One CTA owns one (slice, query-tile):
//   * a producer warp streams the slice into a shared-memory ring with 1-D bulk async copies
//     (cp.async.bulk, completion on an mbarrier) - 16 KB stages, so the bytes in flight do not depend on
//     registers or occupancy;
//   * 8 consumer warps lift their 64 bytes of each stage into registers (conflict-free LDS.128) and test
//     every code against every query of the tile: XOR + POPC + 3-input MIN per (code, query), one compare
//     and one branch per 8 codes.  With PREFILTER the test uses popc(x_lo | x_hi) <= popc(x_lo) + popc(x_hi),
//     a lower bound of the distance that needs half the POPCs; survivors are re-checked exactly.
//   * the per-query distance threshold tau is shared by ALL CTAs of the launch: every appended code is
//     counted once in a global per-query distance histogram, tau = the smallest d with at least k codes
//     counted at distance <= d (a valid upper bound of the final k-th distance), lowered with atomicMin
//     and re-read by every CTA at each step - so the whole grid learns the threshold together and only
//     ~k*ln(N/k) codes per query ever leave the fast path;
//   * the rare survivors go to the query's candidate buffer; a warp compacts it with a bitonic sort when
//     it fills.  Overflow (ties, first step) rolls the step back and replays it one sub-column at a time.
// Each CTA finally writes its k best per query; merge_topk_kernel folds the n_slices partial lists (and,
// multi-GPU, the all-gathered per-shard lists).  No global atomics, no host round trips.
#pragma once
#include "common.cuh"

namespace vc {

constexpr int kScanThreads = 256;                          // consumer threads
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanCtaThreads = kScanThreads + 32;         // + one producer warp
constexpr int kScanU4PerThread = 4;                        // 64 bytes per consumer thread per step
constexpr int kScanStageBytes = kScanThreads * kScanU4PerThread * 16;   // 16 KB
constexpr int kScanSub = 128;                              // threads appending at once on the careful path
constexpr int kScanMaxStages = 8;

struct ScanParams {
  const uint4* codes;        // shard codes, [n][W] u64, 16-byte aligned, padded
  uint64_t n;                // codes in the shard
  uint32_t first_id;         // global id of code 0
  uint32_t id_stride;        // id of code j = first_id + j * id_stride (1, or the number of interleaved shards)
  const uint32_t* queries;   // [nq][2W] u32
  uint32_t nq, k;
  uint32_t QT;               // queries per CTA
  uint32_t BUF;              // candidate-buffer entries per query: power of two >= k + kScanSub
  uint32_t compact_at;       // compact a buffer once it holds this many entries
  uint32_t stages;           // ring depth
  uint32_t n_qtiles, n_slices;
  uint64_t slice_codes;      // codes per slice (multiple of the step size)
  uint32_t interleave;       // 1: slice j = steps j, j + n_slices, ... (all CTAs stream one narrow address window)
  uint64_t* partial;         // [n_slices][nq][k]
  uint32_t* gtau;            // [nq]     launch-wide distance thresholds (initialised to kInfDist)
  uint32_t* ghist;           // [nq][HB] launch-wide distance histograms of appended codes (initialised to 0)
};

template <int W> struct ScanCfg {
  static constexpr int C = 2 * kScanU4PerThread / W;        // codes per thread per step (8 / 4 / 2)
  static constexpr int STEP = kScanThreads * C;             // codes per CTA step (2048 / 1024 / 512)
  static constexpr int QSTRIDE = (2 * W + 1 + 3) / 4 * 4;   // u32 per query record: 2W query words, tau, pad
  static constexpr int HB = 64 * W + 32;                    // histogram bins per query (distances 0..64W)
};

// ---- PTX: mbarrier + 1-D bulk async copy (TMA engine, no tensor map needed) --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// barrier over the consumer threads only (id 1), with an OR-reduction of a predicate
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kScanThreads) : "memory"); }
__device__ __forceinline__ uint32_t consumer_sync_or(uint32_t pred) {
  uint32_t r;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, 1, %2, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
               : "=r"(r) : "r"(pred), "n"(kScanThreads) : "memory");
  return r;
}

// ---- shared-memory layout -----------------------------------------------------------------------------
struct ScanSmem {
  unsigned char* ring;  // [stages][kScanStageBytes]
  uint64_t* buf;        // [QT][BUF]
  uint32_t* qrec;       // [QT][QSTRIDE]  query words, then the distance threshold tau
  uint64_t* tau_key;    // [QT] k-th best packed word at the last compaction
  uint32_t* cnt;        // [QT]
  uint32_t* cnt0;       // [QT] count at step start (roll-back point)
  uint32_t* flag;       // [QT] 1 = overflowed this step / needs the careful path
  uint64_t* full;       // [stages]
  uint64_t* empty;      // [stages]
};

__host__ __device__ inline size_t scan_smem_bytes(uint32_t QT, uint32_t BUF, int qstride, uint32_t stages) {
  return (size_t)stages * kScanStageBytes + (size_t)QT * BUF * 8 + (size_t)QT * qstride * 4 + (size_t)QT * 8 +
         (size_t)QT * 12 + 16 + (size_t)2 * kScanMaxStages * 8;
}

__device__ __forceinline__ ScanSmem scan_carve(unsigned char* base, uint32_t QT, uint32_t BUF, int qstride, uint32_t stages) {
  ScanSmem s;                                   // every section stays 16-byte aligned
  s.ring = base;
  s.buf = (uint64_t*)(base + (size_t)stages * kScanStageBytes);
  s.qrec = (uint32_t*)(s.buf + (size_t)QT * BUF);
  s.tau_key = (uint64_t*)(s.qrec + (size_t)QT * qstride);
  s.cnt = (uint32_t*)(s.tau_key + QT);
  s.cnt0 = s.cnt + QT;
  s.flag = s.cnt0 + QT;
  s.full = (uint64_t*)(((uintptr_t)(s.flag + QT) + 15) & ~(uintptr_t)15);
  s.empty = s.full + kScanMaxStages;
  return s;
}

// Lowers the launch-wide threshold of one query from its global histogram: tau = smallest d with at
// least k codes counted at distance <= d.  Concurrent increments may be missed (counts only grow), which
// can only make the result larger, i.e. still a valid bound.
__device__ __forceinline__ void scan_update_tau(const uint32_t* gh, int hb, uint32_t k, uint32_t lim, uint32_t* gtau_q, uint32_t* stau_q) {
  uint32_t c = 0;
  for (uint32_t j = 0; j < lim; j += 4) {            // 16-byte L2 reads, independent of each other
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(gh + j));
    const uint32_t e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      c += e[t];
      if (c >= k && j + t < lim) { atomicMin(gtau_q, j + t); atomicMin(stau_q, j + t); return; }
    }
  }
}

// Rare path of the fast loop: a code whose exact distance d passed the threshold.  Appends it to the
// query's buffer (flagging overflow), counts it in the launch-wide histogram and tightens tau.
__device__ __noinline__ void scan_append(unsigned char* smem, uint32_t QT, uint32_t BUF, uint32_t stages, int qstride, int hb,
                                         uint32_t tau_off, uint32_t k, uint32_t q, uint32_t d, uint32_t id,
                                         uint32_t* gtau_q, uint32_t* gh) {
  const ScanSmem s = scan_carve(smem, QT, BUF, qstride, stages);
  const uint64_t key = pack_key(d, id);
  if (key >= s.tau_key[q]) return;
  const uint32_t slot = atomicAdd(&s.cnt[q], 1u);
  if (slot < BUF) s.buf[(size_t)q * BUF + slot] = key;
  else s.flag[q] = 1;
  atomicAdd(&gh[d], 1u);
  uint32_t* stau_q = s.qrec + q * qstride + tau_off;
  const uint32_t tau = *(volatile uint32_t*)stau_q;
  if (d < tau) scan_update_tau(gh, hb, k, min(tau, (uint32_t)hb), gtau_q, stau_q);   // only a code strictly inside can lower it
}

template <int W, bool PREFILTER>
__global__ void __launch_bounds__(kScanCtaThreads) scan_topk_kernel(const ScanParams p) {
  using Cfg = ScanCfg<W>;
  constexpr int C = Cfg::C, QS = Cfg::QSTRIDE, HB = Cfg::HB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const ScanSmem s = scan_carve(smem_raw, p.QT, p.BUF, QS, p.stages);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t qtile = blockIdx.x % p.n_qtiles, slice = blockIdx.x / p.n_qtiles;
  const uint32_t q0 = qtile * p.QT;
  const uint32_t nq_here = min(p.QT, p.nq - q0);
  const uint32_t BUF = p.BUF, S = p.stages;
  // contiguous: steps [slice*spp, ...) of the shard; interleaved: steps slice, slice + n_slices, ...
  const uint64_t total_steps = (p.n + Cfg::STEP - 1) / Cfg::STEP;
  const uint64_t spp = p.slice_codes / Cfg::STEP;
  const uint64_t step0 = p.interleave ? slice : (uint64_t)slice * spp;
  const uint64_t stride = p.interleave ? p.n_slices : 1;
  const uint32_t n_steps = p.interleave ? (uint32_t)(total_steps > slice ? (total_steps - slice + p.n_slices - 1) / p.n_slices : 0)
                                        : (uint32_t)(total_steps > step0 ? min(spp, total_steps - step0) : 0);
  const uint64_t end = p.n;

  // ---- set-up: query tile, counters, barriers ------------------------------------------------------
  for (uint32_t i = tid; i < nq_here * QS; i += kScanCtaThreads) {
    const uint32_t q = i / QS, j = i % QS;
    s.qrec[i] = j < 2 * W ? p.queries[(size_t)(q0 + q) * 2 * W + j] : (j == 2 * W ? kInfDist : 0u);
  }
  for (uint32_t q = tid; q < nq_here; q += kScanCtaThreads) {
    s.tau_key[q] = kEmptyKey; s.cnt[q] = 0; s.cnt0[q] = 0; s.flag[q] = 1;   // first step: careful path
  }
  if (tid == 0) {
    for (uint32_t i = 0; i < S; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], kScanWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kScanWarps) {
    // ===== producer warp: one lane feeds the ring =====================================================
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      for (uint32_t i = 0; i < n_steps; ++i, st = (st + 1 == S ? 0 : st + 1), ph ^= (st == 0)) {
        mbar_wait(&s.empty[st], ph ^ 1);
        const uint64_t base = (step0 + (uint64_t)i * stride) * Cfg::STEP;
        const uint64_t codes_here = min((uint64_t)Cfg::STEP, end - base);
        const uint32_t bytes = (uint32_t)((codes_here * W * 8 + 15) & ~15ull);
        mbar_arrive_expect_tx(&s.full[st], bytes);
        bulk_g2s(s.ring + (size_t)st * kScanStageBytes, p.codes + base * W / 2, bytes, &s.full[st]);
      }
    }
    return;
  }

  // ===== consumer warps ==================================================================================
  uint32_t careful = 1;                 // CTA-uniform: this step must take the careful path
  uint32_t first_settle = 1;            // the first settle seeds the launch-wide histogram with the kept entries
  uint32_t st = 0, ph = 0;
  for (uint32_t i = 0; i < n_steps; ++i, st = (st + 1 == S ? 0 : st + 1), ph ^= (st == 0)) {
    const uint64_t base = (step0 + (uint64_t)i * stride) * Cfg::STEP;
    CodeRegs<W> code[C];
    // pick up the launch-wide thresholds the other CTAs have reached (one L2 read per query, off the critical path)
    uint32_t gt = kInfDist;
    if (tid < nq_here) gt = __ldcg(p.gtau + q0 + tid);
    mbar_wait(&s.full[st], ph);
    {
      const uint4* stage = reinterpret_cast<const uint4*>(s.ring + (size_t)st * kScanStageBytes);
#pragma unroll
      for (int u = 0; u < kScanU4PerThread; ++u) {
        if constexpr (W == 1) {          // unit = two codes: local index 2*(u*T + t) + h
          const uint4 v = stage[u * kScanThreads + tid];
          code[2 * u].w[0] = v.x; code[2 * u].w[1] = v.y;
          code[2 * u + 1].w[0] = v.z; code[2 * u + 1].w[1] = v.w;
        } else if constexpr (W == 2) {   // unit = one code: local index u*T + t
          const uint4 v = stage[u * kScanThreads + tid];
          code[u].w[0] = v.x; code[u].w[1] = v.y; code[u].w[2] = v.z; code[u].w[3] = v.w;
        } else {                         // W == 4: code c = units 2*(c*T + t), +1: local index c*T + t
          const int cc = u / 2, h = u % 2;
          const uint4 v = stage[2 * (cc * kScanThreads + tid) + h];
          code[cc].w[4 * h + 0] = v.x; code[cc].w[4 * h + 1] = v.y; code[cc].w[4 * h + 2] = v.z; code[cc].w[4 * h + 3] = v.w;
        }
      }
    }
    // The stage is refilled by cp.async.bulk (async proxy) as soon as all consumer warps have released it, and these loads are
    // generic-proxy reads that may still be in flight when a release issued right behind them lands: measured on 256-bit codes,
    // k = 1000, two CTAs per SM - the other CTA's bitonic sorts keep the shared-memory pipe busy -, a quarter of the scans
    // returned a few wrong distances (tools/flake_probe.py).  So: a proxy fence behind the loads, and the release only after the
    // step's barrier below, when every consumer thread has used its registers (~1 % of the HBM-bound scan's bandwidth).
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (tid < nq_here && gt < s.qrec[tid * QS + 2 * W]) atomicMin(&s.qrec[tid * QS + 2 * W], gt);
    auto local_of = [&](int c) -> uint32_t {
      if constexpr (W == 1) return 2 * ((c / 2) * kScanThreads + tid) + (c & 1);
      else return c * kScanThreads + tid;
    };

    uint32_t appended = 0;
    if (!careful) {
      // ---- fast path: every query of the tile against the C codes in registers ---------------------
#pragma unroll 1
      for (uint32_t q = 0; q < nq_here; ++q) {
        const QRec<W> cur = load_qrec<W, QS>(s.qrec, q);
        const uint32_t* qw = cur.qw;
        const uint32_t tau = cur.tau;
        uint32_t m[C];
        uint32_t mn = 0xFFFFFFFFu;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          m[c] = PREFILTER ? hamming_lower_bound<W>(code[c].w, qw) : hamming_exact<W>(code[c].w, qw);
          mn = min(mn, m[c]);
        }
        if (mn <= tau) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            if (m[c] <= tau) {
              const uint32_t d = PREFILTER ? hamming_exact<W>(code[c].w, qw) : m[c];
              const uint64_t idx = base + local_of(c);
              if (d <= tau && idx < end) {
                scan_append(smem_raw, p.QT, BUF, S, QS, HB, 2 * W, p.k, q, d, p.first_id + (uint32_t)idx * p.id_stride,
                            p.gtau + q0 + q, p.ghist + (size_t)(q0 + q) * HB);
                appended = 1;
              }
            }
          }
        }
      }
    }
    // one barrier per step; it also tells every thread whether anything was appended
    const uint32_t any = consumer_sync_or(appended | careful);
    if (lane == 0) mbar_arrive(&s.empty[st]);          // the stage may be refilled from here on
    if (!any) continue;

    // ---- step epilogue ---------------------------------------------------------------------------------
    uint32_t need_careful = careful;
    if (!careful) {
      uint32_t f = 0;
      for (uint32_t q = tid; q < nq_here; q += kScanThreads) f |= s.flag[q];
      need_careful = consumer_sync_or(f);
    }
    if (need_careful) {
      // overflowed queries (all of them on the first step): roll the step back, then replay it one
      // sub-column (kScanSub threads x one code) at a time with room guaranteed before each
      for (uint32_t q = tid; q < nq_here; q += kScanThreads)
        if (s.flag[q]) s.cnt[q] = s.cnt0[q];
      consumer_sync();
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const uint64_t idx = base + local_of(c);
        const bool valid = idx < end;
        const uint32_t id = p.first_id + (uint32_t)idx * p.id_stride;
        for (uint32_t half = 0; half < kScanThreads / kScanSub; ++half) {
          for (uint32_t q = warp; q < nq_here; q += kScanWarps) {
            if (s.flag[q] && s.cnt[q] + kScanSub > BUF) {
              const uint64_t tk = topk_compact(s.buf + (size_t)q * BUF, &s.cnt[q], BUF, p.k, lane, 32, WarpSync());
              if (lane == 0) {
                s.tau_key[q] = tk;
                if (tk != kEmptyKey) atomicMin(&s.qrec[q * QS + 2 * W], (uint32_t)(tk >> 32));
              }
            }
          }
          consumer_sync();
          if (tid / kScanSub == half) {
            for (uint32_t q = 0; q < nq_here; ++q) {
              if (!s.flag[q]) continue;
              const uint32_t* qr = s.qrec + q * QS;
              uint32_t d = 0;
#pragma unroll
              for (int j = 0; j < 2 * W; ++j) d += __popc(code[c].w[j] ^ qr[j]);
              const uint64_t key = pack_key(d, id);
              if (valid && d <= qr[2 * W] && key < s.tau_key[q]) {
                const uint32_t slot = atomicAdd(&s.cnt[q], 1u);
                s.buf[(size_t)q * BUF + slot] = key;      // slot < BUF by construction
              }
            }
          }
          consumer_sync();
        }
      }
      // settle the replayed queries: compact; on the very first step also count the kept entries in the
      // launch-wide histogram (every code is counted there exactly once: here, or when the fast path appends it)
      for (uint32_t q = warp; q < nq_here; q += kScanWarps) {
        if (!s.flag[q]) continue;
        const uint64_t tk = topk_compact(s.buf + (size_t)q * BUF, &s.cnt[q], BUF, p.k, lane, 32, WarpSync());
        if (first_settle) {
          uint32_t* gh = p.ghist + (size_t)(q0 + q) * HB;
          const uint32_t kept = s.cnt[q];
          for (uint32_t e = lane; e < kept; e += 32) atomicAdd(&gh[(uint32_t)(s.buf[(size_t)q * BUF + e] >> 32)], 1u);
          __syncwarp();
          if (lane == 0) scan_update_tau(gh, HB, p.k, HB, p.gtau + q0 + q, &s.qrec[q * QS + 2 * W]);
        }
        if (lane == 0) {
          s.tau_key[q] = tk;
          if (tk != kEmptyKey) atomicMin(&s.qrec[q * QS + 2 * W], (uint32_t)(tk >> 32));
          s.flag[q] = 0;
        }
      }
      first_settle = 0;
    }
    for (uint32_t q = warp; q < nq_here; q += kScanWarps) {
      if (s.cnt[q] >= p.compact_at) {
        const uint64_t tk = topk_compact(s.buf + (size_t)q * BUF, &s.cnt[q], BUF, p.k, lane, 32, WarpSync());
        if (lane == 0) {
          s.tau_key[q] = tk;
          if (tk != kEmptyKey) atomicMin(&s.qrec[q * QS + 2 * W], (uint32_t)(tk >> 32));
        }
      }
      if (lane == 0) s.cnt0[q] = s.cnt[q];
    }
    careful = 0;
    consumer_sync();
  }

  // ---- final compaction and partial-list write -------------------------------------------------------
  for (uint32_t q = warp; q < nq_here; q += kScanWarps) {
    topk_compact(s.buf + (size_t)q * BUF, &s.cnt[q], BUF, p.k, lane, 32, WarpSync());
    const uint32_t kept = s.cnt[q];
    uint64_t* out = p.partial + ((size_t)slice * p.nq + q0 + q) * p.k;
    for (uint32_t i = lane; i < p.k; i += 32) out[i] = i < kept ? s.buf[(size_t)q * BUF + i] : kEmptyKey;
  }
}

// launch-wide thresholds / histograms of one search call
__global__ void scan_init_kernel(uint32_t* gtau, uint32_t* ghist, uint32_t nq, uint32_t hb) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (uint64_t)nq * hb; i += (uint64_t)gridDim.x * blockDim.x) {
    ghist[i] = 0;
    if (i < nq) gtau[i] = kInfDist;
  }
}

// ---------------------------------------------------------------------------------------------
// merge_topk_kernel: out[q] = k smallest of lists[0..n_lists)[q][0..k), ascending.
// One CTA per query; the same buffer + compaction scheme as above, block-wide.
// ---------------------------------------------------------------------------------------------
constexpr int kMergeThreads = 256;

// grid = (nq, groups): group g folds lists [g*fanin, min((g+1)*fanin, n_lists)) into out[g][q][0..k)
// list_stride = keys between list l and list l + 1 (nq * k when the lists are contiguous)
__global__ void __launch_bounds__(kMergeThreads) merge_topk_kernel(const uint64_t* __restrict__ lists, uint32_t n_lists, uint32_t fanin,
                                                                   uint32_t nq, uint32_t k, uint32_t BUF,
                                                                   uint64_t* __restrict__ out, uint64_t list_stride) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* buf = (uint64_t*)smem_raw;            // [BUF]
  __shared__ uint32_t cnt;
  __shared__ uint64_t tau_s;
  constexpr uint32_t NB = 64 * 4 + 2;             // distance bins 0 .. 256, one more for anything above (topk_compact_block clamps)
  __shared__ uint32_t mh[NB + 2];
  const uint32_t q = blockIdx.x, grp = blockIdx.y, tid = threadIdx.x;
  const uint32_t l0 = grp * fanin, nl = min(fanin, n_lists - l0);
  if (tid == 0) { cnt = 0; tau_s = kEmptyKey; }
  __syncthreads();
  const uint64_t total = (uint64_t)nl * k;
  for (uint64_t base = 0; base < total; base += kMergeThreads) {
    if (cnt + kMergeThreads > BUF) {     // uniform: cnt is stable between barriers
      uint64_t tau = topk_compact_block(buf, &cnt, BUF, k, mh, NB, tid, kMergeThreads);
      if (tid == 0) tau_s = tau;
      __syncthreads();
    }
    const uint64_t i = base + tid;
    if (i < total) {
      const uint64_t l = l0 + i / k, j = i % k;
      const uint64_t key = lists[l * list_stride + (uint64_t)q * k + j];
      if (key < tau_s) { uint32_t slot = atomicAdd(&cnt, 1u); buf[slot] = key; }   // kEmptyKey never passes
    }
    __syncthreads();
  }
  topk_compact_block(buf, &cnt, BUF, k, mh, NB, tid, kMergeThreads);
  const uint32_t kept = cnt;
  uint64_t* o = out + ((size_t)grp * nq + q) * k;
  for (uint32_t i = tid; i < k; i += kMergeThreads) o[i] = i < kept ? buf[i] : kEmptyKey;
}

// keys [nq][k] -> ids / dists / counts
__global__ void unpack_keys_kernel(const uint64_t* __restrict__ keys, uint32_t nq, uint32_t k,
                                   uint32_t* __restrict__ ids, uint32_t* __restrict__ dists, uint32_t* __restrict__ counts) {
  const uint32_t q = blockIdx.x;
  __shared__ uint32_t c;
  if (threadIdx.x == 0) c = 0;
  __syncthreads();
  uint32_t mine = 0;
  for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
    const uint64_t key = keys[(size_t)q * k + i];
    const bool ok = key != kEmptyKey;
    if (ids) ids[(size_t)q * k + i] = ok ? (uint32_t)key : 0xFFFFFFFFu;
    if (dists) dists[(size_t)q * k + i] = ok ? (uint32_t)(key >> 32) : 0xFFFFFFFFu;
    mine += ok;
  }
  if (mine) atomicAdd(&c, mine);
  __syncthreads();
  if (threadIdx.x == 0 && counts) counts[q] = c;
}

}  // namespace vc
