// Index build (K1, K8): substring tables as CSR arrays in HBM.
//
// Replaces load_binarycode of the reference (src/build_hash_tables.cc:25-73: for every code,
// key = binaryToInt(substring t) (Pilaf/image_tools.h:12-18, unsigned), then a network
// get / append {id, code} / put of the whole bucket) by, per table:
//   extract keys -> stable LSD radix sort of (key, ordinal) -> row_ptr -> gather {id, code} payload
// so that a bucket is a contiguous run of the payload arrays, members in ascending id order (the
// order the reference's inserts produce).  For s = 32 a dense row_ptr would need 2^32 entries, so
// the table keeps the reference's occupancy bitmap (src/bitmap.cc:22-38, one bit per bucket index,
// src/generate_bitmap.cc:99-125) plus a rank directory (set bits before each 256-bit block) that
// maps an occupied bucket to its slot in a compact start array.
#pragma once
#include "common.cuh"

namespace vc {

struct TableDev {
  uint32_t* row_ptr;    // dense: [2^s + 1] bucket starts; sparse: [n_unique + 1] starts of occupied buckets
  uint32_t* bitmap;     // sparse only: 2^32 occupancy bits
  uint32_t* rank_dir;   // sparse only: [2^32 / 256] set bits before each 256-bit block
  uint32_t* ids;        // [n]   global ids in bucket order
  uint64_t* codes;      // [n][W] full codes in bucket order (the reference stores {id, code} per bucket member)
  uint32_t n_unique;    // occupied buckets
  uint32_t sparse;      // 1 = bitmap + rank directory
};

constexpr int kBuildThreads = 256;
constexpr int kSortItemsPerThread = 16;
constexpr int kSortChunk = kBuildThreads * kSortItemsPerThread;     // 4096 elements per block
constexpr int kSortWarpRange = kSortChunk / (kBuildThreads / 32);   // 512 contiguous elements per warp
constexpr uint32_t kRankBlockBits = 256;                              // one 32-byte sector of bitmap per rank block

// key of table t: bits [t*s, (t+1)*s) of the little-endian code
template <int W>
__global__ void extract_keys_kernel(const uint64_t* __restrict__ codes, uint64_t n, uint32_t table, uint32_t sbits,
                                    uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t off = table * sbits, word = off >> 6, sh = off & 63;
  const uint64_t mask = sbits == 32 ? 0xFFFFFFFFull : ((1ull << sbits) - 1);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    keys[i] = (uint32_t)((codes[i * W + word] >> sh) & mask);
    vals[i] = (uint32_t)i;
  }
}

// ---- exclusive scan over u32 (three-phase, recursive on the block sums) ---------------------------
constexpr int kScanTile = 2048;   // elements per block: 256 threads x 8

__global__ void __launch_bounds__(kBuildThreads) scan_reduce_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                                                     uint32_t* __restrict__ sums) {
  __shared__ uint32_t wsum[kBuildThreads / 32];
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t s = 0;
  for (int j = 0; j < kScanTile / kBuildThreads; ++j) {
    const uint64_t i = base + (uint64_t)j * kBuildThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kBuildThreads / 32; ++w) t += wsum[w];
    sums[blockIdx.x] = t;
  }
}

// out[i] = base[block] + exclusive prefix inside the tile; in and out may alias
__global__ void __launch_bounds__(kBuildThreads) scan_apply_kernel(const uint32_t* in, uint64_t n,
                                                                    const uint32_t* __restrict__ block_base, uint32_t* out) {
  __shared__ uint32_t wsum[kBuildThreads / 32];
  constexpr int IPT = kScanTile / kBuildThreads;   // 8 consecutive items per thread
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * IPT;
  uint32_t v[IPT], tsum = 0;
#pragma unroll
  for (int j = 0; j < IPT; ++j) { v[j] = base + j < n ? in[base + j] : 0u; tsum += v[j]; }
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = tsum;
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += y; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
  for (uint32_t w = 0; w < warp; ++w) wbase += wsum[w];
  uint32_t run = (block_base ? block_base[blockIdx.x] : 0u) + wbase + incl - tsum;
#pragma unroll
  for (int j = 0; j < IPT; ++j) { if (base + j < n) out[base + j] = run; run += v[j]; }
}

// ---- stable LSD radix sort, 8-bit digit, (key, val) pairs -------------------------------------------
// hist[d * nblocks + b] = number of keys of block b with digit d
__global__ void __launch_bounds__(kBuildThreads) radix_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, uint32_t shift,
                                                                    uint32_t nblocks, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kSortChunk;
  for (int j = 0; j < kSortItemsPerThread; ++j) {
    const uint64_t i = base + (uint64_t)j * kBuildThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// offs = exclusive scan of hist (digit-major), so offs[d * nblocks + b] is where block b's digit-d run starts
__global__ void __launch_bounds__(kBuildThreads) radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                       uint64_t n, uint32_t shift, uint32_t nblocks,
                                                                       const uint32_t* __restrict__ offs,
                                                                       uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  constexpr int NW = kBuildThreads / 32;
  __shared__ uint32_t whist[NW][256];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = 0; w < NW; ++w) whist[w][threadIdx.x] = 0;
  __syncthreads();
  const uint64_t wbase = (uint64_t)blockIdx.x * kSortChunk + (uint64_t)warp * kSortWarpRange;
  // phase 1: per-warp digit histogram of its contiguous range
  for (int r = 0; r < kSortWarpRange / 32; ++r) {
    const uint64_t i = wbase + (uint64_t)r * 32 + lane;
    if (i < n) atomicAdd(&whist[warp][(keys_in[i] >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  // phase 2: digit d (= thread): start of each warp's run inside the block's run
  {
    uint32_t run = offs[(size_t)threadIdx.x * nblocks + blockIdx.x];
    for (int w = 0; w < NW; ++w) { const uint32_t c = whist[w][threadIdx.x]; whist[w][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
  // phase 3: stable ranking, 32 elements at a time, in order
  for (int r = 0; r < kSortWarpRange / 32; ++r) {
    const uint64_t i = wbase + (uint64_t)r * 32 + lane;
    const bool valid = i < n;
    const uint32_t active = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t key = keys_in[i], val = vals_in[i];
      const uint32_t d = (key >> shift) & 0xffu;
      const uint32_t peers = __match_any_sync(active, d);
      const uint32_t rank = __popc(peers & ((1u << lane) - 1));
      const uint32_t dst = whist[warp][d] + rank;
      keys_out[dst] = key;
      vals_out[dst] = val;
      __syncwarp(active);
      if (rank == 0) whist[warp][d] += __popc(peers);
    }
    __syncwarp();
  }
}

// ---- CSR row_ptr from the sorted keys (dense tables) -------------------------------------------------
// thread j in [0, n]: fills row_ptr for every key in (sk[j-1], sk[j]] with j (sk[-1] = -1, sk[n] = n_buckets)
__global__ void row_ptr_kernel(const uint32_t* __restrict__ sk, uint64_t n, uint64_t n_buckets, uint32_t* __restrict__ row_ptr) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= n; j += (uint64_t)gridDim.x * blockDim.x) {
    const int64_t lo = j == 0 ? -1 : (int64_t)sk[j - 1];
    const int64_t hi = j == n ? (int64_t)n_buckets : (int64_t)sk[j];
    for (int64_t key = lo + 1; key <= hi; ++key) row_ptr[key] = (uint32_t)j;
  }
}

// ---- sparse tables (s = 32): occupancy bitmap, rank directory, compact starts -------------------------
__global__ void bitmap_set_kernel(const uint32_t* __restrict__ sk, uint64_t n, uint32_t* __restrict__ bitmap) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t key = sk[j];
    if (j == 0 || sk[j - 1] != key) atomicOr(&bitmap[key >> 5], 1u << (key & 31));   // src/bitmap.cc:28-32 set_idx
  }
}
// counts[b] = set bits of 256-bit block b
__global__ void bitmap_block_count_kernel(const uint32_t* __restrict__ bitmap, uint64_t n_blocks, uint32_t* __restrict__ counts) {
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (uint64_t)gridDim.x * blockDim.x) {
    const uint4* p = reinterpret_cast<const uint4*>(bitmap + b * 8);
    const uint4 x = p[0], y = p[1];
    counts[b] = __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w) + __popc(y.x) + __popc(y.y) + __popc(y.z) + __popc(y.w);
  }
}
// rank of an occupied bucket = set bits below it
__device__ __forceinline__ uint32_t sparse_rank(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ rank_dir, uint32_t key) {
  const uint32_t blk = key >> 8, w = (key >> 5) & 7, bit = key & 31;
  uint32_t r = rank_dir[blk];
  const uint32_t* p = bitmap + (size_t)blk * 8;
  for (uint32_t i = 0; i < w; ++i) r += __popc(p[i]);
  return r + __popc(p[w] & ((1u << bit) - 1));
}
__global__ void sparse_starts_kernel(const uint32_t* __restrict__ sk, uint64_t n, const uint32_t* __restrict__ bitmap,
                                     const uint32_t* __restrict__ rank_dir, uint32_t n_unique, uint32_t* __restrict__ starts) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= n; j += (uint64_t)gridDim.x * blockDim.x) {
    if (j == n) { starts[n_unique] = (uint32_t)n; continue; }
    const uint32_t key = sk[j];
    if (j == 0 || sk[j - 1] != key) starts[sparse_rank(bitmap, rank_dir, key)] = (uint32_t)j;
  }
}

// bucket lookup shared by the MIH kernel and vc_bucket_get
__device__ __forceinline__ void table_lookup(const TableDev& t, uint32_t key, uint32_t& start, uint32_t& len) {
  if (!t.sparse) {
    start = t.row_ptr[key];
    len = t.row_ptr[key + 1] - start;
  } else {
    if ((t.bitmap[key >> 5] >> (key & 31)) & 1u) {      // src/bitmap.cc:22-26 get_idx
      const uint32_t r = sparse_rank(t.bitmap, t.rank_dir, key);
      start = t.row_ptr[r];
      len = t.row_ptr[r + 1] - start;
    } else { start = 0; len = 0; }
  }
}
__global__ void bucket_lookup_kernel(TableDev t, uint32_t key, uint32_t* out /*[2]*/) {
  uint32_t s, l;
  table_lookup(t, key, s, l);
  out[0] = s; out[1] = l;
}
// occupancy bitmap of a dense table, reference layout (bit i of word i/32)
__global__ void dense_bitmap_kernel(const uint32_t* __restrict__ row_ptr, uint64_t n_words, uint32_t* __restrict__ words) {
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t v = 0;
    for (uint32_t b = 0; b < 32; ++b) v |= (row_ptr[w * 32 + b + 1] > row_ptr[w * 32 + b]) ? (1u << b) : 0u;
    words[w] = v;
  }
}

// payload in bucket order: ids[j] = first_id + perm[j] * id_stride, codes[j] = main[perm[j]]
template <int W>
__global__ void gather_payload_kernel(const uint64_t* __restrict__ main_codes, const uint32_t* __restrict__ perm, uint64_t n,
                                      uint32_t first_id, uint32_t id_stride, uint32_t* __restrict__ ids, uint64_t* __restrict__ codes) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t o = perm[j];
    ids[j] = first_id + o * id_stride;
#pragma unroll
    for (int w = 0; w < W; ++w) codes[j * W + w] = main_codes[(uint64_t)o * W + w];
  }
}

// A table read back from a file (vc_index_load) before any kernel trusts it: row_ptr non-decreasing and ending at n; for a
// sparse table also popcount(bitmap) == n_unique and rank_dir = exclusive prefix sums of the 256-bit block counts.
// bad[0] is raised on any violation (a truncated or corrupt file would otherwise mean out-of-bounds reads in every search).
__global__ void validate_row_ptr_kernel(const uint32_t* __restrict__ row_ptr, uint64_t entries, uint64_t n, uint32_t* bad) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < entries; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v = row_ptr[i];
    if (v > n || (i + 1 < entries && row_ptr[i + 1] < v) || (i == 0 && v != 0) || (i + 1 == entries && v != n)) atomicOr(bad, 1u);
  }
}
__global__ void validate_bitmap_kernel(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ rank_dir, uint64_t n_blocks,
                                       unsigned long long* total, uint32_t* bad) {
  // rank_dir[b] must equal the set bits before block b: checked locally (rank_dir[b + 1] - rank_dir[b] == bits of block b)
  unsigned long long mine = 0;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c = 0;
    for (uint32_t w = 0; w < kRankBlockBits / 32; ++w) c += __popc(bitmap[b * (kRankBlockBits / 32) + w]);
    mine += c;
    if (b == 0 && rank_dir[0] != 0) atomicOr(bad, 1u);
    if (b + 1 < n_blocks && rank_dir[b + 1] - rank_dir[b] != c) atomicOr(bad, 1u);
  }
  if (mine) atomicAdd(total, mine);
}

template <int W>
__global__ void synth_codes_kernel(uint64_t* __restrict__ codes, uint64_t n, uint64_t first_id, uint64_t id_stride, uint64_t seed) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * W; i += (uint64_t)gridDim.x * blockDim.x)
    codes[i] = synth_word(seed, first_id + (i / W) * id_stride, (uint32_t)(i % W));
}

}  // namespace vc
