// Shared device helpers of the verticut_b200 kernels (sm_100a).
//
//  - packed result word  (dist << 32 | id)           src/search_worker.cc:12-13,254-256
//  - Hamming distance    XOR + POPC                   Pilaf/image_tools.h:21-33
//  - TopK buffer in shared memory with warp / block bitonic compaction; replaces the
//    std::priority_queue of src/search_worker.cc:160,192-197 and src/linear_search.cc:43,51-56
//    with the canonical tie rule (ascending packed word = ascending (dist, id)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vc {

constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kInfDist = 0x7FFFFFFFu;   // "no threshold yet"

__host__ __device__ __forceinline__ uint64_t pack_key(uint32_t dist, uint32_t id) {
  return ((uint64_t)dist << 32) | id;
}

// 128-bit streaming load: read-only path, do not allocate in L1 (each code is used once per CTA).
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// ---------------------------------------------------------------------------------------------
// Bitonic sort of P (power of two) u64 keys in shared memory by `nthreads` cooperating threads
// (tid in [0, nthreads)).  SYNC() must synchronise exactly those threads.
// ---------------------------------------------------------------------------------------------
template <class SyncFn>
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* a, uint32_t P, uint32_t tid, uint32_t nthreads, SyncFn sync) {
  for (uint32_t k2 = 2; k2 <= P; k2 <<= 1) {
    for (uint32_t j = k2 >> 1; j > 0; j >>= 1) {
      for (uint32_t t = tid; t < (P >> 1); t += nthreads) {
        uint32_t l = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        uint32_t r = l | j;
        bool asc = (l & k2) == 0;
        uint64_t x = a[l], y = a[r];
        if ((x > y) == asc) { a[l] = y; a[r] = x; }
      }
      sync();
    }
  }
}

__device__ __forceinline__ uint32_t next_pow2(uint32_t v) {
  return v <= 1 ? 1u : 1u << (32 - __clz(v - 1));
}

// ---------------------------------------------------------------------------------------------
// One query's candidate buffer: `buf` holds `*cnt` unsorted packed words (cnt may have run past
// `cap` if appends overflowed; those were dropped by the caller).  compact() sorts, keeps the k
// smallest, and returns the new threshold: the k-th smallest word when the buffer holds k, else
// kEmptyKey (everything passes).  Executed by one warp (all 32 lanes) or by a whole block.
// ---------------------------------------------------------------------------------------------
template <class SyncFn>
__device__ __forceinline__ uint64_t topk_compact(uint64_t* buf, uint32_t* cnt, uint32_t cap, uint32_t k,
                                                 uint32_t tid, uint32_t nthreads, SyncFn sync) {
  uint32_t n = min(*cnt, cap);
  uint32_t P = max(next_pow2(n), 2u);
  for (uint32_t i = n + tid; i < P; i += nthreads) buf[i] = kEmptyKey;
  sync();
  bitonic_sort_smem(buf, P, tid, nthreads, sync);
  uint32_t kept = min(n, k);
  uint64_t tau = (kept == k) ? buf[k - 1] : kEmptyKey;
  sync();
  if (tid == 0) *cnt = kept;
  sync();
  return tau;
}

struct WarpSync  { __device__ __forceinline__ void operator()() const { __syncwarp(); } };
struct BlockSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

// Exclusive prefix sums of `per` consecutive bins per lane, over one warp: lane l owns bins [l * per, (l + 1) * per).
// Returns the lane's own sum in `mine` and the sum of all lower lanes' bins as the result.
__device__ __forceinline__ uint32_t warp_bins_excl(const uint32_t* bins, uint32_t nb, uint32_t per, uint32_t lane, uint32_t& mine) {
  uint32_t s = 0;
  for (uint32_t j = 0; j < per; ++j) { const uint32_t b = lane * per + j; if (b < nb) s += bins[b]; }
  uint32_t incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += v; }
  mine = s;
  return incl - s;
}

// ---------------------------------------------------------------------------------------------
// topk_compact for a whole block (nthreads a multiple of 32, all threads call) when the buffer is much longer than k:
// the bitonic sort costs P log^2 P shared-memory traffic, and most of a long buffer cannot be among the k smallest.  So
// first a histogram of the distance fields (bin = min(dist, nb - 1)), the first bin b_k at which k entries are reached,
// and an in-place compaction to the entries of bins <= b_k - k plus the ties of the k-th distance; only those are sorted.
// Same result as topk_compact (the k smallest, ascending, and the k-th word).  `hist` = nb + 2 words of shared memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t topk_compact_block(uint64_t* buf, uint32_t* cnt, uint32_t cap, uint32_t k,
                                                       uint32_t* hist, uint32_t nb, uint32_t tid, uint32_t nthreads) {
  const uint32_t n = min(*cnt, cap);
  if (n > 256 && n >= 2 * k + 64) {                                   // uniform over the block
    for (uint32_t i = tid; i < nb + 2; i += nthreads) hist[i] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += nthreads) {            // one atomic per warp and distinct bin: the entries crowd into 3 - 4 bins
      const uint32_t i = base + tid;
      const uint32_t bin = i < n ? min((uint32_t)(buf[i] >> 32), nb - 1) : 0xFFFFFFFFu;
      const uint32_t peers = __match_any_sync(0xffffffffu, bin);
      if (i < n && (tid & 31) == (uint32_t)__ffs(peers) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (tid < 32) {
      const uint32_t per = (nb + 31) / 32;
      uint32_t mine;
      const uint32_t excl = warp_bins_excl(hist, nb, per, tid, mine);
      if (excl < k && excl + mine >= k) {                              // exactly one lane: the block holds n >= k entries
        uint32_t cum = excl, b = tid * per;
        for (;; ++b) { cum += hist[b]; if (cum >= k) break; }
        hist[nb] = b;
      }
    }
    __syncthreads();
    const uint32_t bk = hist[nb];
    // in place: chunk c reads entries [c T, (c + 1) T) before the barrier and writes after it to slots below the number
    // kept so far, which is at most (c + 1) T - entries no later chunk still has to read
    for (uint32_t base = 0; base < n; base += nthreads) {
      const uint32_t i = base + tid;
      const uint64_t key = i < n ? buf[i] : kEmptyKey;
      const bool keep = i < n && min((uint32_t)(key >> 32), nb - 1) <= bk;
      __syncthreads();
      const uint32_t mask = __ballot_sync(0xffffffffu, keep);
      uint32_t wbase = 0;
      if ((tid & 31) == 0 && mask) wbase = atomicAdd(&hist[nb + 1], (uint32_t)__popc(mask));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (keep) buf[wbase + __popc(mask & ((1u << (tid & 31)) - 1u))] = key;
    }
    __syncthreads();
    if (tid == 0) *cnt = hist[nb + 1];
    __syncthreads();
  }
  return topk_compact(buf, cnt, cap, k, tid, nthreads, BlockSync());
}

// Hamming distance helpers on 64-bit words held as two 32-bit halves.
template <int W> struct CodeRegs { uint32_t w[2 * W]; };   // W 64-bit words = 2W 32-bit words

template <int W>
__device__ __forceinline__ uint32_t hamming_exact(const uint32_t* __restrict__ c, const uint32_t* __restrict__ q) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 2 * W; ++i) d += __popc(c[i] ^ q[i]);
  return d;
}
// Lower bound of the distance with half the POPCs: popc(x | y) <= popc(x) + popc(y).
// Two LOP3 per 64 bits: t = a ^ b, then (c ^ d) | t as ONE three-input LOP3 (LUT 0xBE).  Written in PTX because the
// compiler otherwise keeps both XORs as separate instructions to share them with the exact path (3 LOP3 per test: the ALU
// pipe, not the POPC pipe, then bounds the loop).
__device__ __forceinline__ uint32_t xor_or(uint32_t a, uint32_t b, uint32_t t) {
  uint32_t y;
  asm("lop3.b32 %0, %1, %2, %3, 0xBE;" : "=r"(y) : "r"(a), "r"(b), "r"(t));
  return y;
}
template <int W>
__device__ __forceinline__ uint32_t hamming_lower_bound(const uint32_t* __restrict__ c, const uint32_t* __restrict__ q) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) d += __popc(xor_or(c[2 * i + 1], q[2 * i + 1], c[2 * i] ^ q[2 * i]));
  return d;
}

// One staged query record in shared memory: 2W query words, then the distance threshold (stride QS u32,
// 16-byte aligned).  Loaded with 128-bit LDS; the loops prefetch record q+1 while testing record q.
template <int W> struct QRec { uint32_t qw[2 * W]; uint32_t tau; };
template <int W, int QS>
__device__ __forceinline__ QRec<W> load_qrec(const uint32_t* base, uint32_t q) {
  QRec<W> r;
  const uint4* qv = reinterpret_cast<const uint4*>(base + q * QS);
  if constexpr (W == 1) {
    const uint4 v = qv[0];
    r.qw[0] = v.x; r.qw[1] = v.y; r.tau = v.z;
  } else {
#pragma unroll
    for (int j = 0; j < W / 2; ++j) {
      const uint4 v = qv[j];
      r.qw[4 * j] = v.x; r.qw[4 * j + 1] = v.y; r.qw[4 * j + 2] = v.z; r.qw[4 * j + 3] = v.w;
    }
    r.tau = base[q * QS + 2 * W];
  }
  return r;
}

// splitmix64-based synthetic code words (must match oracle/verticut_oracle.c: vo_synth_word)
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t synth_word(uint64_t seed, uint64_t id, uint32_t word) {
  return splitmix64(splitmix64(seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(word + 1))) + id);
}

}  // namespace vc
