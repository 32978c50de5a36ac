// Linear scan (K6) + top-k selection (K5) + partial-list merge (K7).
//
// Replaces search_K_nearest_neighbors(int k) of the reference, src/linear_search.cc:39-64
// (for every code: compute_hamming_dist, max-heap of k) for a BATCH of queries, with the
// canonical tie rule (k smallest packed words dist<<32|id, ascending).
//
// Decomposition: the database shard is cut into `n_slices` contiguous slices, the query batch into
// `n_qtiles` tiles of QT queries.  One CTA owns one (slice, query-tile): it keeps the QT queries,
// their thresholds and their candidate buffers in shared memory, streams the slice through
// registers (4 x 128-bit loads per thread per step) and tests every code against every query of
// the tile.  Codes whose distance does not beat the query's current k-th best (the common case)
// cost XOR + POPC + MIN only; the rare survivors are appended to the query's shared-memory buffer,
// which a warp compacts with a bitonic sort when it fills.  Each CTA finally writes its k best
// per query; merge_topk_kernel folds the n_slices partial lists (and, multi-GPU, the all-gathered
// per-shard lists) into the final answer.  No global atomics, no host round trips.
#pragma once
#include "common.cuh"

namespace vc {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanU4PerThread = 4;                       // 64 bytes per thread per step

struct ScanParams {
  const uint4* codes;        // shard codes, [n][W] u64, 16-byte aligned
  uint64_t n;                // codes in the shard
  uint32_t first_id;         // global id of code 0
  const uint32_t* queries;   // [nq][2W] u32
  uint32_t nq, k;
  uint32_t QT;               // queries per CTA
  uint32_t BUF;              // candidate-buffer entries per query: power of two >= k + kScanThreads
  uint32_t compact_at;       // compact a buffer once it holds this many entries
  uint32_t n_qtiles, n_slices;
  uint64_t slice_codes;      // codes per slice (multiple of the step size)
  uint64_t* partial;         // [n_slices][nq][k]
};

template <int W> struct ScanCfg {
  static constexpr int C = 2 * kScanU4PerThread / W;        // codes per thread per step (8 / 4 / 2)
  static constexpr int STEP = kScanThreads * C;             // codes per CTA step (2048 / 1024 / 512)
  static constexpr int QSTRIDE = (2 * W + 1 + 3) / 4 * 4;   // u32 per query record: 2W query words, tau, pad
};

struct ScanSmem {
  uint64_t* buf;      // [QT][BUF]
  uint64_t* tau_key;  // [QT]
  uint32_t* qrec;     // [QT][QSTRIDE]  query words then the distance threshold
  uint32_t* cnt;      // [QT]
  uint32_t* cnt0;     // [QT] count at step start (roll-back point)
  uint32_t* ovf;      // [QT] 1 = needs the careful path this step
  uint32_t* any_ovf;  // [1]
};

__host__ __device__ inline size_t scan_smem_bytes(uint32_t QT, uint32_t BUF, int qstride) {
  return (size_t)QT * BUF * 8 + (size_t)QT * 8 + (size_t)QT * qstride * 4 + (size_t)QT * 12 + 16;
}

__device__ __forceinline__ ScanSmem scan_carve(unsigned char* base, uint32_t QT, uint32_t BUF, int qstride) {
  ScanSmem s;                                   // every section stays 16-byte aligned for uint4 reads of qrec
  s.buf = (uint64_t*)base;
  s.qrec = (uint32_t*)(s.buf + (size_t)QT * BUF);
  s.tau_key = (uint64_t*)(s.qrec + (size_t)QT * qstride);
  s.cnt = (uint32_t*)(s.tau_key + QT);
  s.cnt0 = s.cnt + QT;
  s.ovf = s.cnt0 + QT;
  s.any_ovf = s.ovf + QT;
  return s;
}

// Rare path of the fast loop: exact distance, exact (dist,id) test, append or flag overflow.
template <int W>
__device__ __noinline__ void scan_append(const ScanSmem s, uint32_t BUF, int qstride, uint32_t q, CodeRegs<W> c, uint32_t id) {
  const uint32_t* qr = s.qrec + q * qstride;
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 2 * W; ++i) d += __popc(c.w[i] ^ qr[i]);
  uint64_t key = pack_key(d, id);
  if (key < s.tau_key[q]) {
    uint32_t slot = atomicAdd(&s.cnt[q], 1u);
    if (slot < BUF) s.buf[(size_t)q * BUF + slot] = key;
    else { s.ovf[q] = 1; *s.any_ovf = 1; }
  }
}

template <int W, bool PREFILTER>
__global__ void __launch_bounds__(kScanThreads) scan_topk_kernel(const ScanParams p) {
  using Cfg = ScanCfg<W>;
  constexpr int C = Cfg::C, QS = Cfg::QSTRIDE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ScanSmem s = scan_carve(smem_raw, p.QT, p.BUF, QS);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t qtile = blockIdx.x % p.n_qtiles, slice = blockIdx.x / p.n_qtiles;
  const uint32_t q0 = qtile * p.QT;
  const uint32_t nq_here = min(p.QT, p.nq - q0);
  const uint32_t BUF = p.BUF;

  // ---- stage the query tile --------------------------------------------------------------
  for (uint32_t i = tid; i < nq_here * QS; i += kScanThreads) {
    uint32_t q = i / QS, j = i % QS;
    s.qrec[i] = j < 2 * W ? p.queries[(size_t)(q0 + q) * 2 * W + j] : (j == 2 * W ? kInfDist : 0u);
  }
  for (uint32_t q = tid; q < nq_here; q += kScanThreads) {
    s.tau_key[q] = kEmptyKey; s.cnt[q] = 0; s.cnt0[q] = 0; s.ovf[q] = 1;   // first step: careful path
  }
  if (tid == 0) *s.any_ovf = 1;
  __syncthreads();

  const uint64_t beg = (uint64_t)slice * p.slice_codes;
  const uint64_t end = min(p.n, beg + p.slice_codes);

  for (uint64_t base = beg; base < end; base += Cfg::STEP) {
    // ---- load this step's codes: unit u of thread t = uint4 (base*W/2 + u*T + t) ------------
    CodeRegs<W> code[C];
    uint32_t local[C];            // index of the code inside the step
    {
      const uint64_t u4_base = base * W / 2;               // 16-byte units (base is a STEP multiple)
      const uint64_t u4_end = (p.n * W + 1) / 2;           // units that exist (last may be half-valid when W=1)
#pragma unroll
      for (int u = 0; u < kScanU4PerThread; ++u) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if constexpr (W == 1) {          // one unit = two codes
          const uint64_t idx = u4_base + (uint64_t)u * kScanThreads + tid;
          if (idx < u4_end) v = ld_stream_u4(p.codes + idx);
          code[2 * u].w[0] = v.x; code[2 * u].w[1] = v.y;
          code[2 * u + 1].w[0] = v.z; code[2 * u + 1].w[1] = v.w;
          local[2 * u] = 2 * (u * kScanThreads + tid);
          local[2 * u + 1] = 2 * (u * kScanThreads + tid) + 1;
        } else if constexpr (W == 2) {   // one unit = one code
          const uint64_t idx = u4_base + (uint64_t)u * kScanThreads + tid;
          if (idx < u4_end) v = ld_stream_u4(p.codes + idx);
          code[u].w[0] = v.x; code[u].w[1] = v.y; code[u].w[2] = v.z; code[u].w[3] = v.w;
          local[u] = u * kScanThreads + tid;
        } else {                         // W == 4: two consecutive units = one code
          const int cc = u / 2, h = u % 2;
          const uint64_t idx = u4_base + 2 * ((uint64_t)cc * kScanThreads + tid) + h;
          if (idx < u4_end) v = ld_stream_u4(p.codes + idx);
          code[cc].w[4 * h + 0] = v.x; code[cc].w[4 * h + 1] = v.y; code[cc].w[4 * h + 2] = v.z; code[cc].w[4 * h + 3] = v.w;
          local[cc] = cc * kScanThreads + tid;
        }
      }
    }

    if (*s.any_ovf == 0) {
      // ---- fast path: every query of the tile against the C codes in registers --------------
#pragma unroll 1
      for (uint32_t q = 0; q < nq_here; ++q) {
        uint32_t qw[2 * W];
        const uint4* qv = reinterpret_cast<const uint4*>(s.qrec + q * QS);
        uint32_t tau;
        if constexpr (W == 1) {
          const uint4 r = qv[0];
          qw[0] = r.x; qw[1] = r.y; tau = r.z;
        } else {
#pragma unroll
          for (int i = 0; i < W / 2; ++i) {
            const uint4 r = qv[i];
            qw[4 * i] = r.x; qw[4 * i + 1] = r.y; qw[4 * i + 2] = r.z; qw[4 * i + 3] = r.w;
          }
          tau = s.qrec[q * QS + 2 * W];
        }
        uint32_t m[C];
        uint32_t mn = 0xFFFFFFFFu;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          m[c] = PREFILTER ? hamming_lower_bound<W>(code[c].w, qw) : hamming_exact<W>(code[c].w, qw);
          mn = min(mn, m[c]);
        }
        if (mn <= tau) {
#pragma unroll
          for (int c = 0; c < C; ++c)
            if (m[c] <= tau && base + local[c] < end)
              scan_append<W>(s, BUF, QS, q, code[c], p.first_id + (uint32_t)(base + local[c]));
        }
      }
    }
    __syncthreads();

    // ---- step epilogue: roll back overflowed queries, compact full buffers ---------------------
    if (*s.any_ovf) {
      // careful path (first step of the CTA, or a buffer overflowed): one code column at a time,
      // with room for a full column guaranteed before each.
      for (uint32_t q = tid; q < nq_here; q += kScanThreads)
        if (s.ovf[q]) s.cnt[q] = s.cnt0[q];
      __syncthreads();
#pragma unroll
      for (int c = 0; c < C; ++c) {
        for (uint32_t q = warp; q < nq_here; q += kScanWarps) {
          if (s.ovf[q] && s.cnt[q] + kScanThreads > BUF) {
            uint64_t tau = topk_compact(s.buf + (size_t)q * BUF, &s.cnt[q], BUF, p.k, lane, 32, WarpSync());
            if (lane == 0) { s.tau_key[q] = tau; s.qrec[q * QS + 2 * W] = tau == kEmptyKey ? kInfDist : (uint32_t)(tau >> 32); }
          }
        }
        __syncthreads();
        const bool valid = base + local[c] < end;
        const uint32_t id = p.first_id + (uint32_t)(base + local[c]);
        for (uint32_t q = 0; q < nq_here; ++q) {
          if (!s.ovf[q]) continue;
          const uint32_t* qr = s.qrec + q * QS;
          uint32_t d = 0;
#pragma unroll
          for (int i = 0; i < 2 * W; ++i) d += __popc(code[c].w[i] ^ qr[i]);
          const uint64_t key = pack_key(d, id);
          if (valid && key < s.tau_key[q]) {
            uint32_t slot = atomicAdd(&s.cnt[q], 1u);
            s.buf[(size_t)q * BUF + slot] = key;      // slot < BUF by construction
          }
        }
        __syncthreads();
      }
      for (uint32_t q = tid; q < nq_here; q += kScanThreads) s.ovf[q] = 0;
      if (tid == 0) *s.any_ovf = 0;
      __syncthreads();
    }
    for (uint32_t q = warp; q < nq_here; q += kScanWarps) {
      if (s.cnt[q] >= p.compact_at) {
        uint64_t tau = topk_compact(s.buf + (size_t)q * BUF, &s.cnt[q], BUF, p.k, lane, 32, WarpSync());
        if (lane == 0) { s.tau_key[q] = tau; s.qrec[q * QS + 2 * W] = tau == kEmptyKey ? kInfDist : (uint32_t)(tau >> 32); }
      }
      if (lane == 0) s.cnt0[q] = s.cnt[q];
    }
    __syncthreads();
  }

  // ---- final compaction and partial-list write -------------------------------------------------
  for (uint32_t q = warp; q < nq_here; q += kScanWarps) {
    topk_compact(s.buf + (size_t)q * BUF, &s.cnt[q], BUF, p.k, lane, 32, WarpSync());
    const uint32_t kept = s.cnt[q];
    uint64_t* out = p.partial + ((size_t)slice * p.nq + q0 + q) * p.k;
    for (uint32_t i = lane; i < p.k; i += 32) out[i] = i < kept ? s.buf[(size_t)q * BUF + i] : kEmptyKey;
  }
}

// ---------------------------------------------------------------------------------------------
// merge_topk_kernel: out[q] = k smallest of lists[0..n_lists)[q][0..k), ascending.
// One CTA per query; the same buffer + compaction scheme as above, block-wide.
// ---------------------------------------------------------------------------------------------
constexpr int kMergeThreads = 256;

// grid = (nq, groups): group g folds lists [g*fanin, min((g+1)*fanin, n_lists)) into out[g][q][0..k)
__global__ void __launch_bounds__(kMergeThreads) merge_topk_kernel(const uint64_t* __restrict__ lists, uint32_t n_lists, uint32_t fanin,
                                                                   uint32_t nq, uint32_t k, uint32_t BUF,
                                                                   uint64_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* buf = (uint64_t*)smem_raw;            // [BUF]
  __shared__ uint32_t cnt;
  __shared__ uint64_t tau_s;
  const uint32_t q = blockIdx.x, grp = blockIdx.y, tid = threadIdx.x;
  const uint32_t l0 = grp * fanin, nl = min(fanin, n_lists - l0);
  if (tid == 0) { cnt = 0; tau_s = kEmptyKey; }
  __syncthreads();
  const uint64_t total = (uint64_t)nl * k;
  for (uint64_t base = 0; base < total; base += kMergeThreads) {
    if (cnt + kMergeThreads > BUF) {     // uniform: cnt is stable between barriers
      uint64_t tau = topk_compact(buf, &cnt, BUF, k, tid, kMergeThreads, BlockSync());
      if (tid == 0) tau_s = tau;
      __syncthreads();
    }
    const uint64_t i = base + tid;
    if (i < total) {
      const uint64_t l = l0 + i / k, j = i % k;
      const uint64_t key = lists[(l * nq + q) * k + j];
      if (key < tau_s) { uint32_t slot = atomicAdd(&cnt, 1u); buf[slot] = key; }   // kEmptyKey never passes
    }
    __syncthreads();
  }
  topk_compact(buf, &cnt, BUF, k, tid, kMergeThreads, BlockSync());
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
