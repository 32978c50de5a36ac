// C ABI of libverticut_gpu.so (include/verticut_gpu.h): index object in HBM, build, search, merge.
// Host code only orchestrates: every byte of the hot path is touched by the kernels in scan.cuh,
// build.cuh and mih.cuh.  There is no CPU fallback - without a CUDA device the calls fail.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/verticut_gpu.h"
#include "build.cuh"
#include "mih.cuh"
#include "scan.cuh"
#include "bmih.cuh"
#include "tcverify.cuh"
#include "tcverify_v1.cuh"
#include "xchg.cuh"

using namespace vc;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(e_ == cudaErrorMemoryAllocation ? VC_ERR_NOMEM : VC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
  } while (0)

// ------------------------------------------------------------------------------------------------
// small RAII helpers
// ------------------------------------------------------------------------------------------------
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// grow-only device / pinned-host scratch
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return VC_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = std::max(bytes, (size_t)256);
    CU(cudaMalloc(&p, want));
    cap = want;
    return VC_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return VC_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = std::max(bytes, (size_t)256);
    CU(cudaMallocHost(&p, want));
    cap = want;
    return VC_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// ------------------------------------------------------------------------------------------------
// the index object
// ------------------------------------------------------------------------------------------------
struct vc_index {
  int device = 0;
  uint32_t bits = 0, W = 0, m = 0, sbits = 0, first_id = 0;
  uint32_t id_stride = 1;         // id of code j = first_id + j * id_stride: > 1 for one of G interleaved shards ("id_stride", set before adding codes)
  uint64_t n = 0, cap = 0;        // codes held / capacity (codes)
  uint64_t* d_codes = nullptr;    // main table id -> code (src/linear_search.cc:45-46), [cap][W]
  bool built = false;
  TableDev tab[kMaxTables];       // host copy of the table descriptors
  TableDev* d_tab = nullptr;      // device copy
  uint64_t table_bytes = 0;
  int num_sms = 0;
  size_t smem_optin = 0;
  // scratch
  DevBuf d_q, d_partial, d_partial2, d_keys, d_ids, d_dists, d_counts, d_stats, d_small, d_gstate;
  DevBuf b_state, b_buckets, b_qlist, b_items, b_redo, b_idh, b_approx;   // batched MIH
  PinBuf h_q, h_ids, h_dists, h_counts, h_stats, h_small;
  // knobs
  int64_t scan_prefilter = -1;    // -1 auto, 0 off, 1 on
  int64_t scan_qt = 0;            // 0 auto
  int64_t scan_waves = 0;         // 0 auto
  int64_t scan_stages = 0;        // 0 auto: ring depth
  int64_t scan_batched = -1;        // large batches through the verify kernel: -1 auto, 0 never, 1 always
  int64_t scan_batched_min = 2;     // measured on 1 B x 64-bit codes: batch 2 / 4 reach 0.93 / 0.87 of the HBM peak through the verify kernel, 0.74 / 0.63 through the ring kernel
  int64_t last_scan_batched = 0;
  int64_t scan_ctas_per_sm = 2;    // shared-memory plan targets this many resident CTAs per SM
  int64_t scan_interleave = 1;    // slices interleaved step-wise (1) or contiguous (0)
  int64_t scan_smem_kb = 0;       // 0 auto: shared memory budget per CTA for the query tile
  int64_t merge_fanin = 64;
  int64_t mih_batched = -1;       // -1 auto, 0 per-query kernel only, 1 bucket-stationary batched path whenever legal
  int64_t mih_prefilter = -1;
  int64_t mih_cpi_steps = 0;
  int64_t mih_wide = -1;
  int64_t mih_min_bucket = 64;
  int64_t host_chunk = 32768;              // knob "host.chunk": queries per device batch of the host-buffer calls
  uint32_t max_bucket_len = 0;    // longest bucket of the dense tables (0: not computed yet for this build)
  int64_t mih_boot_sample = 0;    // codes per query of the threshold bootstrap (0: max(16384, 16 k))
  // a step that probes radius 0 together with radius 1 verifies the queries' own (radius-0) buckets first, on the exact distance, and
  // the other buckets after them with the lower-bound filter: -1 = when the step has >= 2 queries per probed bucket, 0 never, 1 always.
  // Measured slower than one launch on the exact distance (batch 16384: 10.6 vs 9.9 ms; profiles/ab_r02.md): off.
  int64_t mih_r0_first = 0;
  // speculative thresholds of the batched exact search (bmih_decide_kernel): the largest final k-th distance of the previous batch
  // on this index caps every query's starting threshold; a query it was too small for is detected and redone, exactly.  Measured
  // (profiles/spec_r02.log): 1 B codes, batch 16 384: first step 9.96 -> 7.23 ms, the search 51.5 -> 48.7 ms.  Learned from and
  // applied to LARGE batches only (>= kSpecMinBatch queries): a handful of queries says little about the next batch's largest k-th
  // distance, and a guess that is too small sends its misses through the per-query kernel
  int64_t mih_speculate = 1;      // knob "mih.speculate": 0 off, 1 batches of >= 1024 queries, 2 every batch (tests)
  int64_t mih_spec_force = -1;    // knob "mih.spec_tau": >= 0 forces this guess (tests), -1 learns it
  bool spec_valid = false;
  uint32_t spec_tau = 0, spec_k = 0;
  int64_t last_spec_tau = -1, last_spec_fail = 0;
  int64_t mih_split_r0 = 0;       // 1: radius 0 is a search step of its own (tighter thresholds for radius 1) instead of being probed together with radius 1
  int64_t mih_cap = 0;            // candidate-buffer entries per query (0: bmih_cap_for(k)); small values force the overflow path in tests
  int64_t mih_global_key = 1;     // id-sharded search: exchange a bound on the k-th key of the whole database before table-granular steps
  // tensor-core verify kernel (tcverify.cuh): 0 never (default: measured slower than the POPC kernels),
  // 1 whenever legal, -1 by size (scan: >= scan.tc_min queries; MIH: steps with >= mih.tc_ratio queries per code)
  // scan: -1 by default = the first version of the kernel (tcverify_v1.cuh, = 2) for 64-bit codes and >= scan.tc_min queries, where it
  // measured 1.13 - 1.16 x the POPC kernel (profiles/tc_r02.md); MIH: 0 (a probed bucket is shared by too few queries to fill a tile)
  int64_t scan_tc = -1, scan_tc_min = 256, last_scan_tc = 0;
  int64_t mih_tc = 0, mih_tc_ratio = 10, last_mih_tc_steps = 0;
  int64_t tc_trace = 0;           // debug: dump CTA 0's pipeline timestamps of the last tensor-core launch to stderr
  DevBuf b_trace;
  int64_t tc_units = 0, tc_flagged = 0, tc_hits = 0;     // statistics of the tensor-core launches of the last search
  vc_allreduce_fn allreduce_fn = nullptr;
  void* allreduce_user = nullptr;
  // peer exchange over NVLink peer memory (xchg.cuh): window of this rank, the peers' windows, running exchange number
  uint32_t x_rank = 0, x_world = 0;
  uint64_t x_slot_bytes = 0;
  unsigned char* x_window = nullptr;
  unsigned char* x_peer[kXchgMaxWorld] = {nullptr};
  bool x_ipc[kXchgMaxWorld] = {false};
  bool x_open = false;
  uint32_t x_seq = 0, x_done_total = 0;
  uint32_t* x_ctr = nullptr;      // device: [0] CTAs that finished a push (running total), [1] a wait timed out
  DevBuf x_local;                 // local top-k of a sharded search call
  bool x_fuse_finish = false;     // the next batched search pushes its result rows from bmih_finish_kernel itself ...
  bool x_pushed = false;          // ... and says here whether it did
  XchgDev x_finish;               // ... into this exchange
  int64_t x_enabled = 1;          // knob "xchg": 0 = keep the all-reduce hook / NCCL for everything
  // knob "xchg.allreduce": the histogram sums over peer memory?  1 always, 0 never (hook), -1 = only while a rank's payload times the
  // number of peers stays below 2 MB.  Every rank stores its whole payload into every window (G - 1 times the payload out of each
  // GPU), where the NVSwitch reduces an NCCL all-reduce in the fabric: measured at 8 GPUs, 6.3 MB per rank, 0.10 ms per exchange
  // slower than ncclAllReduce (profiles/scale_r02.json); the result rows (an all-gather, no reduction to be had) always go over
  // peer memory.
  int64_t x_allreduce = -1;
  // knob "xchg.emulate" = G > 1 (measurements only, tools/probe.py shards=G): this index behaves like ONE of G id-shards of a
  // G times larger database - every cross-shard sum is replaced by "times G", statistically what G shards of uniform codes give -
  // so that the per-shard kernels of a multi-GPU search can be timed and profiled on one GPU.  The answers are not those of any database.
  int64_t x_emulate = 0;
  int64_t last_xchg = 0;          // exchanges of the last search that went over peer memory
  bool sharded() const { return allreduce_fn != nullptr || x_open || x_emulate > 1; }
  int64_t mih_table_steps = -1;   // stop rule tested after every table of a radius: 0 never (reference-like radius steps), 1 always, -1 auto    // batched path when the average bucket holds at least this many codes
  int64_t last_mih_batched = 0, last_mih_levels = 0, last_mih_items = 0, last_mih_bucket_codes = 0, last_mih_redo = 0;
  // optional device-side timing of the dominant kernel of the last search ("profile" = 1)
  int64_t profile = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;
  cudaEvent_t lev[2 * 34] = {nullptr};   // batched MIH: one (start, stop) pair per radius level around the verify kernel
  cudaEvent_t lev_x[2 * 34] = {nullptr}; // ... and (step begin, step end): probes + work items before the verify kernel, settle + exchange + decide after it
  cudaEvent_t ev_s0 = nullptr, ev_s1 = nullptr;   // whole batched search (first launch .. finish kernel)
  int lev_used = 0;                        // > 0: last search was batched, sum these pairs
  std::vector<int64_t> step_exec;     // ... and the tests the verify kernel really executed
  std::vector<int64_t> step_codes, step_pairs;   // per step of the last batched search: codes of distinct probed buckets, code-query tests
  // counters
  int64_t launches = 0;           // kernels launched by this index since creation
  int64_t last_scan_grid = 0, last_scan_qt = 0, last_scan_slices = 0, last_scan_smem = 0, last_scan_occ = 0, last_scan_stages = 0;
};

static void free_tables(vc_index* ix) {
  for (uint32_t t = 0; t < kMaxTables; ++t) {
    TableDev& T = ix->tab[t];
    if (T.row_ptr) cudaFree(T.row_ptr);
    if (T.bitmap) cudaFree(T.bitmap);
    if (T.rank_dir) cudaFree(T.rank_dir);
    if (T.ids) cudaFree(T.ids);
    if (T.codes) cudaFree(T.codes);
    memset(&T, 0, sizeof T);
  }
  ix->table_bytes = 0;
  ix->built = false;
}

static int grid_for(uint64_t n, int threads, int num_sms) {
  uint64_t blocks = (n + threads - 1) / threads;
  uint64_t cap = (uint64_t)num_sms * 16;
  return (int)std::max<uint64_t>(1, std::min(blocks, cap));
}

extern "C" {

const char* vc_last_error(void) { return g_err.c_str(); }
int vc_abi_version(void) { return VC_ABI_VERSION; }

int vc_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(VC_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}

uint64_t vc_synth_word(uint64_t seed, uint64_t id, uint32_t word) { return synth_word(seed, id, word); }

int vc_index_create(int device, uint32_t code_bits, uint32_t n_tables, uint32_t first_id, vc_index** out) {
  if (!out) return fail(VC_ERR_ARG, "out is null");
  *out = nullptr;
  if (code_bits != 64 && code_bits != 128 && code_bits != 256)
    return fail(VC_ERR_ARG, "code_bits must be 64, 128 or 256 (got %u)", code_bits);
  uint32_t sbits = 0;
  if (n_tables) {
    if (n_tables > kMaxTables || code_bits % n_tables) return fail(VC_ERR_ARG, "code_bits %u not divisible into %u tables", code_bits, n_tables);
    sbits = code_bits / n_tables;
    if (sbits != 8 && sbits != 16 && sbits != 32)
      return fail(VC_ERR_ARG, "substring width %u unsupported (8, 16 or 32 bits; bucket index is a uint32)", sbits);
  }
  int ndev = vc_device_count();
  if (ndev < 0) return ndev;
  if (device < 0 || device >= ndev) return fail(VC_ERR_CUDA, "CUDA device %d not available (%d visible)", device, ndev);
  DeviceGuard g(device);
  if (!g.ok) return fail(VC_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  vc_index* ix = new vc_index();
  memset(ix->tab, 0, sizeof ix->tab);
  ix->device = device; ix->bits = code_bits; ix->W = code_bits / 64; ix->m = n_tables; ix->sbits = sbits; ix->first_id = first_id;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ix; return fail(VC_ERR_CUDA, "cudaGetDeviceProperties failed"); }
  ix->num_sms = prop.multiProcessorCount;
  ix->smem_optin = prop.sharedMemPerBlockOptin;
  // binomial table for the probe enumerator
  static uint32_t binom[33][33];
  for (int n = 0; n <= 32; ++n)
    for (int k = 0; k <= 32; ++k)
      binom[n][k] = k == 0 ? 1u : (n == 0 ? 0u : (uint32_t)std::min<uint64_t>((uint64_t)binom[n - 1][k - 1] + binom[n - 1][k], 0xFFFFFFFFull));
  cudaError_t e = cudaMemcpyToSymbol(c_binom, binom, sizeof binom);
  if (e != cudaSuccess) { delete ix; return fail(VC_ERR_CUDA, "cudaMemcpyToSymbol: %s", cudaGetErrorString(e)); }
  *out = ix;
  return VC_OK;
}

void vc_index_destroy(vc_index* ix) {
  if (!ix) return;
  DeviceGuard g(ix->device);
  cudaDeviceSynchronize();
  free_tables(ix);
  if (ix->d_tab) cudaFree(ix->d_tab);
  if (ix->d_codes) cudaFree(ix->d_codes);
  if (ix->ev0) { cudaEventDestroy(ix->ev0); cudaEventDestroy(ix->ev1); }
  for (cudaEvent_t e : ix->lev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : ix->lev_x) if (e) cudaEventDestroy(e);
  if (ix->ev_s0) { cudaEventDestroy(ix->ev_s0); cudaEventDestroy(ix->ev_s1); }
  DevBuf* db[] = {&ix->d_q, &ix->d_partial, &ix->d_partial2, &ix->d_keys, &ix->d_ids, &ix->d_dists, &ix->d_counts, &ix->d_stats, &ix->d_small, &ix->d_gstate,
                   &ix->b_state, &ix->b_buckets, &ix->b_qlist, &ix->b_items, &ix->b_redo, &ix->b_idh, &ix->b_approx, &ix->b_trace};
  for (DevBuf* b : db) b->release();
  PinBuf* pb[] = {&ix->h_q, &ix->h_ids, &ix->h_dists, &ix->h_counts, &ix->h_stats, &ix->h_small};
  for (PinBuf* b : pb) b->release();
  for (uint32_t g2 = 0; g2 < ix->x_world; ++g2) if (ix->x_ipc[g2] && ix->x_peer[g2]) cudaIpcCloseMemHandle(ix->x_peer[g2]);
  if (ix->x_window) cudaFree(ix->x_window);
  if (ix->x_ctr) cudaFree(ix->x_ctr);
  ix->x_local.release();
  delete ix;
}

int vc_index_get_info(const vc_index* ix, vc_index_info* info) {
  if (!ix || !info) return fail(VC_ERR_ARG, "null argument");
  info->n_codes = ix->n; info->code_bits = ix->bits; info->n_tables = ix->m; info->substring_bits = ix->sbits;
  info->first_id = ix->first_id; info->device = ix->device; info->built = ix->built ? 1 : 0;
  info->device_bytes = ix->cap * ix->W * 8 + ix->table_bytes;
  return VC_OK;
}

// capacity is kept a multiple of 4096 codes (+ one 4096 pad) so that 128-bit loads of the last codes stay in bounds
static int reserve_codes(vc_index* ix, uint64_t total) {
  if (total && (total - 1) * ix->id_stride > 0xFFFFFFFFull - ix->first_id) return fail(VC_ERR_ARG, "ids would exceed 32 bits (uint32 id, src/image_search.proto:4)");
  if (total <= ix->cap) return VC_OK;
  uint64_t want = std::max(total, ix->cap + ix->cap / 2);
  want = (want + 4095) / 4096 * 4096;
  uint64_t* nd = nullptr;
  CU(cudaMalloc(&nd, (want + 4096) * ix->W * 8));
  if (ix->d_codes) {
    cudaError_t e = cudaMemcpy(nd, ix->d_codes, ix->n * ix->W * 8, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) { cudaFree(nd); return fail(VC_ERR_CUDA, "cudaMemcpy D2D: %s", cudaGetErrorString(e)); }
    cudaFree(ix->d_codes);
  }
  ix->d_codes = nd;
  ix->cap = want;
  return VC_OK;
}

int vc_index_add(vc_index* ix, const void* codes, uint64_t n) {
  if (!ix || (!codes && n)) return fail(VC_ERR_ARG, "null argument");
  if (n == 0) return VC_OK;
  DeviceGuard g(ix->device);
  int rc = reserve_codes(ix, ix->n + n);
  if (rc) return rc;
  CU(cudaMemcpy(ix->d_codes + ix->n * ix->W, codes, n * ix->W * 8, cudaMemcpyHostToDevice));
  ix->n += n;
  ix->built = false;
  return VC_OK;
}

int vc_index_add_device(vc_index* ix, const void* d_codes, uint64_t n) {
  if (!ix || (!d_codes && n)) return fail(VC_ERR_ARG, "null argument");
  if (n == 0) return VC_OK;
  DeviceGuard g(ix->device);
  int rc = reserve_codes(ix, ix->n + n);
  if (rc) return rc;
  CU(cudaMemcpy(ix->d_codes + ix->n * ix->W, d_codes, n * ix->W * 8, cudaMemcpyDeviceToDevice));
  ix->n += n;
  ix->built = false;
  return VC_OK;
}

int vc_index_add_synthetic(vc_index* ix, uint64_t n, uint64_t seed) {
  if (!ix) return fail(VC_ERR_ARG, "null argument");
  if (n == 0) return VC_OK;
  DeviceGuard g(ix->device);
  int rc = reserve_codes(ix, ix->n + n);
  if (rc) return rc;
  uint64_t* dst = ix->d_codes + ix->n * ix->W;
  const uint64_t first = (uint64_t)ix->first_id + ix->n * ix->id_stride;
  const int grid = grid_for(n * ix->W, 256, ix->num_sms);
  if (ix->W == 1) synth_codes_kernel<1><<<grid, 256>>>(dst, n, first, ix->id_stride, seed);
  else if (ix->W == 2) synth_codes_kernel<2><<<grid, 256>>>(dst, n, first, ix->id_stride, seed);
  else synth_codes_kernel<4><<<grid, 256>>>(dst, n, first, ix->id_stride, seed);
  ix->launches++;
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  ix->n += n;
  ix->built = false;
  return VC_OK;
}

// ------------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------------
static int exclusive_scan_inplace(vc_index* ix, uint32_t* d, uint64_t n, cudaStream_t st) {
  // three-phase scan, recursive on block sums; temporary sums are cudaMalloc'ed per level (build-time only)
  const uint64_t nb = (n + kScanTile - 1) / kScanTile;
  if (nb <= 1) {
    scan_apply_kernel<<<1, kBuildThreads, 0, st>>>(d, n, nullptr, d);
    ix->launches++;
    CU(cudaGetLastError());
    return VC_OK;
  }
  uint32_t* sums = nullptr;
  CU(cudaMalloc(&sums, nb * sizeof(uint32_t)));
  scan_reduce_kernel<<<(unsigned)nb, kBuildThreads, 0, st>>>(d, n, sums);
  ix->launches++;
  int rc = exclusive_scan_inplace(ix, sums, nb, st);
  if (rc == VC_OK) {
    scan_apply_kernel<<<(unsigned)nb, kBuildThreads, 0, st>>>(d, n, sums, d);
    ix->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail(VC_ERR_CUDA, "scan_apply: %s", cudaGetErrorString(e));
  }
  cudaStreamSynchronize(st);
  cudaFree(sums);
  return rc;
}

}  // extern "C"

template <int W>
static int build_tables_impl(vc_index* ix) {
  const uint64_t n = ix->n;
  const uint32_t m = ix->m, sbits = ix->sbits;
  cudaStream_t st = 0;
  uint32_t *kA = nullptr, *kB = nullptr, *vA = nullptr, *vB = nullptr, *hist = nullptr;
  const uint64_t nalloc = std::max<uint64_t>(n, 1);
  const uint32_t nblocks = (uint32_t)std::max<uint64_t>(1, (n + kSortChunk - 1) / kSortChunk);
  int rc = VC_OK;
  auto cleanup = [&]() {
    if (kA) cudaFree(kA); if (kB) cudaFree(kB); if (vA) cudaFree(vA); if (vB) cudaFree(vB); if (hist) cudaFree(hist);
  };
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(e_ == cudaErrorMemoryAllocation ? VC_ERR_NOMEM : VC_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
  CUB(cudaMalloc(&kA, nalloc * 4)); CUB(cudaMalloc(&kB, nalloc * 4));
  CUB(cudaMalloc(&vA, nalloc * 4)); CUB(cudaMalloc(&vB, nalloc * 4));
  CUB(cudaMalloc(&hist, (size_t)256 * nblocks * 4));
  const int g1 = grid_for(nalloc, 256, ix->num_sms);
  for (uint32_t t = 0; t < m; ++t) {
    TableDev& T = ix->tab[t];
    extract_keys_kernel<W><<<g1, 256, 0, st>>>(ix->d_codes, n, t, sbits, kA, vA);
    ix->launches++;
    uint32_t *ks = kA, *vs = vA, *kd = kB, *vd = vB;
    for (uint32_t shift = 0; shift < sbits && n > 0; shift += 8) {
      radix_hist_kernel<<<nblocks, kBuildThreads, 0, st>>>(ks, n, shift, nblocks, hist);
      ix->launches++;
      rc = exclusive_scan_inplace(ix, hist, (uint64_t)256 * nblocks, st);
      if (rc) { cleanup(); return rc; }
      radix_scatter_kernel<<<nblocks, kBuildThreads, 0, st>>>(ks, vs, n, shift, nblocks, hist, kd, vd);
      ix->launches++;
      std::swap(ks, kd); std::swap(vs, vd);
    }
    CUB(cudaGetLastError());
    // payload
    CUB(cudaMalloc(&T.ids, (nalloc + 64) * 4));
    CUB(cudaMalloc(&T.codes, (nalloc + 64) * W * 8));
    ix->table_bytes += (nalloc + 64) * (4 + W * 8);
    gather_payload_kernel<W><<<g1, 256, 0, st>>>(ix->d_codes, vs, n, ix->first_id, ix->id_stride, T.ids, T.codes);
    ix->launches++;
    if (sbits <= 16) {
      const uint64_t nbuckets = 1ull << sbits;
      T.sparse = 0;
      CUB(cudaMalloc(&T.row_ptr, (nbuckets + 1) * 4));
      ix->table_bytes += (nbuckets + 1) * 4;
      row_ptr_kernel<<<grid_for(n + 1, 256, ix->num_sms), 256, 0, st>>>(ks, n, nbuckets, T.row_ptr);
      ix->launches++;
    } else {
      // s = 32: occupancy bitmap + rank directory + compact starts
      T.sparse = 1;
      const uint64_t nwords = 1ull << (sbits - 5);
      const uint64_t nrb = (1ull << sbits) / kRankBlockBits;
      CUB(cudaMalloc(&T.bitmap, nwords * 4));
      CUB(cudaMalloc(&T.rank_dir, nrb * 4));
      ix->table_bytes += nwords * 4 + nrb * 4;
      CUB(cudaMemsetAsync(T.bitmap, 0, nwords * 4, st));
      if (n) { bitmap_set_kernel<<<g1, 256, 0, st>>>(ks, n, T.bitmap); ix->launches++; }
      bitmap_block_count_kernel<<<grid_for(nrb, 256, ix->num_sms), 256, 0, st>>>(T.bitmap, nrb, T.rank_dir);
      ix->launches++;
      uint32_t last_cnt = 0, last_rank = 0;
      CUB(cudaMemcpyAsync(&last_cnt, T.rank_dir + nrb - 1, 4, cudaMemcpyDeviceToHost, st));
      CUB(cudaStreamSynchronize(st));
      rc = exclusive_scan_inplace(ix, T.rank_dir, nrb, st);
      if (rc) { cleanup(); return rc; }
      CUB(cudaMemcpyAsync(&last_rank, T.rank_dir + nrb - 1, 4, cudaMemcpyDeviceToHost, st));
      CUB(cudaStreamSynchronize(st));
      T.n_unique = last_rank + last_cnt;
      CUB(cudaMalloc(&T.row_ptr, ((size_t)T.n_unique + 1) * 4));
      ix->table_bytes += ((size_t)T.n_unique + 1) * 4;
      sparse_starts_kernel<<<grid_for(n + 1, 256, ix->num_sms), 256, 0, st>>>(ks, n, T.bitmap, T.rank_dir, T.n_unique, T.row_ptr);
      ix->launches++;
    }
    CUB(cudaGetLastError());
    CUB(cudaStreamSynchronize(st));
  }
  cleanup();
#undef CUB
  if (!ix->d_tab) CU(cudaMalloc(&ix->d_tab, sizeof(TableDev) * kMaxTables));
  CU(cudaMemcpy(ix->d_tab, ix->tab, sizeof(TableDev) * kMaxTables, cudaMemcpyHostToDevice));
  return VC_OK;
}

extern "C" {

int vc_index_build(vc_index* ix) {
  if (!ix) return fail(VC_ERR_ARG, "null argument");
  DeviceGuard g(ix->device);
  free_tables(ix);
  if (ix->m == 0) { ix->built = true; return VC_OK; }
  if (!ix->d_codes) { int rc = reserve_codes(ix, 1); if (rc) return rc; }
  int rc = ix->W == 1 ? build_tables_impl<1>(ix) : ix->W == 2 ? build_tables_impl<2>(ix) : build_tables_impl<4>(ix);
  if (rc) { free_tables(ix); return rc; }
  ix->built = true; ix->max_bucket_len = 0; ix->spec_valid = false;
  return VC_OK;
}

// ------------------------------------------------------------------------------------------------
// persistence
// ------------------------------------------------------------------------------------------------
namespace {
struct FileHeader {
  char magic[8];                 // "VCIX0002" (0001: no id_stride)
  uint32_t bits, m, sbits, first_id;
  uint64_t n;
  uint32_t id_stride, reserved;
};
struct TableHeader { uint32_t sparse, n_unique; uint64_t row_ptr_entries; };
const size_t kIoChunk = (size_t)64 << 20;

int dev_to_file(FILE* f, const void* d, size_t bytes, std::vector<char>& stage) {
  for (size_t off = 0; off < bytes; off += kIoChunk) {
    const size_t nb = std::min(kIoChunk, bytes - off);
    CU(cudaMemcpy(stage.data(), (const char*)d + off, nb, cudaMemcpyDeviceToHost));
    if (fwrite(stage.data(), 1, nb, f) != nb) return fail(VC_ERR_STATE, "short write");
  }
  return VC_OK;
}
int file_to_dev(FILE* f, void* d, size_t bytes, std::vector<char>& stage) {
  for (size_t off = 0; off < bytes; off += kIoChunk) {
    const size_t nb = std::min(kIoChunk, bytes - off);
    if (fread(stage.data(), 1, nb, f) != nb) return fail(VC_ERR_STATE, "short read (truncated index file)");
    CU(cudaMemcpy((char*)d + off, stage.data(), nb, cudaMemcpyHostToDevice));
  }
  return VC_OK;
}
}  // namespace

int vc_index_save(vc_index* ix, const char* path) {
  if (!ix || !path) return fail(VC_ERR_ARG, "null argument");
  if (ix->m && !ix->built) return fail(VC_ERR_STATE, "tables are not built");
  DeviceGuard g(ix->device);
  FILE* f = fopen(path, "wb");
  if (!f) return fail(VC_ERR_ARG, "cannot open %s for writing", path);
  std::vector<char> stage(kIoChunk);
  FileHeader h;
  memcpy(h.magic, "VCIX0002", 8);
  h.bits = ix->bits; h.m = ix->m; h.sbits = ix->sbits; h.first_id = ix->first_id; h.n = ix->n;
  h.id_stride = ix->id_stride; h.reserved = 0;
  int rc = fwrite(&h, sizeof h, 1, f) == 1 ? VC_OK : fail(VC_ERR_STATE, "short write");
  if (!rc && ix->n) rc = dev_to_file(f, ix->d_codes, ix->n * ix->W * 8, stage);
  for (uint32_t t = 0; t < ix->m && !rc; ++t) {
    const TableDev& T = ix->tab[t];
    TableHeader th;
    th.sparse = T.sparse; th.n_unique = T.n_unique;
    th.row_ptr_entries = T.sparse ? (uint64_t)T.n_unique + 1 : (1ull << ix->sbits) + 1;
    if (fwrite(&th, sizeof th, 1, f) != 1) { rc = fail(VC_ERR_STATE, "short write"); break; }
    rc = dev_to_file(f, T.row_ptr, th.row_ptr_entries * 4, stage);
    if (!rc && T.sparse) rc = dev_to_file(f, T.bitmap, (1ull << (ix->sbits - 5)) * 4, stage);
    if (!rc && T.sparse) rc = dev_to_file(f, T.rank_dir, ((1ull << ix->sbits) / kRankBlockBits) * 4, stage);
    if (!rc && ix->n) rc = dev_to_file(f, T.ids, ix->n * 4, stage);
    if (!rc && ix->n) rc = dev_to_file(f, T.codes, ix->n * ix->W * 8, stage);
  }
  if (fclose(f) != 0 && !rc) rc = fail(VC_ERR_STATE, "close failed");
  return rc;
}

int vc_index_load(int device, const char* path, vc_index** out) {
  if (!path || !out) return fail(VC_ERR_ARG, "null argument");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return fail(VC_ERR_ARG, "cannot open %s", path);
  FileHeader h;
  if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "VCIX0002", 8) != 0) { fclose(f); return fail(VC_ERR_ARG, "%s is not a verticut index file", path); }
  vc_index* ix = nullptr;
  int rc = vc_index_create(device, h.bits, h.m, h.first_id, &ix);
  if (rc) { fclose(f); return rc; }
  if (ix->sbits != h.sbits || h.id_stride == 0) { fclose(f); vc_index_destroy(ix); return fail(VC_ERR_ARG, "corrupt header"); }
  ix->id_stride = h.id_stride;
  DeviceGuard g(device);
  std::vector<char> stage(kIoChunk);
  rc = reserve_codes(ix, std::max<uint64_t>(h.n, 1));
  if (!rc && h.n) rc = file_to_dev(f, ix->d_codes, h.n * ix->W * 8, stage);
  ix->n = h.n;
  auto dmalloc = [&](void** p, size_t bytes) -> int { CU(cudaMalloc(p, bytes)); ix->table_bytes += bytes; return VC_OK; };
  const uint64_t nalloc = std::max<uint64_t>(h.n, 1);
  for (uint32_t t = 0; t < h.m && !rc; ++t) {
    TableDev& T = ix->tab[t];
    TableHeader th;
    if (fread(&th, sizeof th, 1, f) != 1) { rc = fail(VC_ERR_STATE, "short read (truncated index file)"); break; }
    const uint64_t expect = th.sparse ? (uint64_t)th.n_unique + 1 : (1ull << h.sbits) + 1;
    if (th.row_ptr_entries != expect || (th.sparse != 0) != (h.sbits > 16)) { rc = fail(VC_ERR_ARG, "corrupt table header"); break; }
    T.sparse = th.sparse; T.n_unique = th.n_unique;
    if ((rc = dmalloc((void**)&T.row_ptr, th.row_ptr_entries * 4))) break;
    if ((rc = file_to_dev(f, T.row_ptr, th.row_ptr_entries * 4, stage))) break;
    if (T.sparse) {
      const uint64_t nwords = 1ull << (h.sbits - 5), nrb = (1ull << h.sbits) / kRankBlockBits;
      if ((rc = dmalloc((void**)&T.bitmap, nwords * 4)) || (rc = file_to_dev(f, T.bitmap, nwords * 4, stage))) break;
      if ((rc = dmalloc((void**)&T.rank_dir, nrb * 4)) || (rc = file_to_dev(f, T.rank_dir, nrb * 4, stage))) break;
    }
    if ((rc = dmalloc((void**)&T.ids, (nalloc + 64) * 4)) || (rc = dmalloc((void**)&T.codes, (nalloc + 64) * ix->W * 8))) break;
    if (h.n && ((rc = file_to_dev(f, T.ids, h.n * 4, stage)) || (rc = file_to_dev(f, T.codes, h.n * ix->W * 8, stage)))) break;
  }
  fclose(f);
  if (!rc && h.m) {
    cudaError_t e = cudaMalloc(&ix->d_tab, sizeof(TableDev) * kMaxTables);
    if (e == cudaSuccess) e = cudaMemcpy(ix->d_tab, ix->tab, sizeof(TableDev) * kMaxTables, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = fail(VC_ERR_CUDA, "table descriptors: %s", cudaGetErrorString(e));
  }
  if (!rc && h.m) {
    // nothing of the file is trusted before it has been checked on the device
    uint32_t* d_bad = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_bad, 16);
    if (e == cudaSuccess) e = cudaMemset(d_bad, 0, 16);
    for (uint32_t t = 0; t < h.m && e == cudaSuccess; ++t) {
      const TableDev& T = ix->tab[t];
      const uint64_t entries = T.sparse ? (uint64_t)T.n_unique + 1 : (1ull << h.sbits) + 1;
      validate_row_ptr_kernel<<<grid_for(entries, 256, ix->num_sms), 256>>>(T.row_ptr, entries, h.n, d_bad);
      if (T.sparse) {
        const uint64_t nrb = (1ull << h.sbits) / kRankBlockBits;
        CU(cudaMemset(d_bad + 2, 0, 8));
        validate_bitmap_kernel<<<grid_for(nrb, 256, ix->num_sms), 256>>>(T.bitmap, T.rank_dir, nrb, (unsigned long long*)(d_bad + 2), d_bad);
        unsigned long long bits = 0;
        e = cudaMemcpy(&bits, d_bad + 2, 8, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && bits != T.n_unique) rc = fail(VC_ERR_ARG, "corrupt index file: table %u has %llu occupied buckets, header says %u", t, bits, T.n_unique);
      }
      ix->launches += T.sparse ? 2 : 1;
    }
    uint32_t bad = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
    if (d_bad) cudaFree(d_bad);
    if (e != cudaSuccess) rc = fail(VC_ERR_CUDA, "index validation: %s", cudaGetErrorString(e));
    else if (!rc && bad) rc = fail(VC_ERR_ARG, "corrupt index file: a bucket directory is not a non-decreasing sequence ending at the code count");
  }
  if (rc) { vc_index_destroy(ix); return rc; }
  ix->built = true; ix->max_bucket_len = 0; ix->spec_valid = false;
  *out = ix;
  return VC_OK;
}

// ------------------------------------------------------------------------------------------------
// BaseProxy::get
// ------------------------------------------------------------------------------------------------
int vc_bucket_get(vc_index* ix, uint32_t table, uint32_t index, uint32_t* ids, void* codes, uint32_t cap, uint32_t* n_out) {
  if (!ix || !n_out) return fail(VC_ERR_ARG, "null argument");
  *n_out = 0;
  if (!ix->built || ix->m == 0) return fail(VC_ERR_STATE, "tables are not built");
  if (table >= ix->m) return fail(VC_ERR_ARG, "table %u out of range (%u tables)", table, ix->m);
  if (ix->sbits < 32 && index >= (1u << ix->sbits)) return VC_NOT_FOUND;
  DeviceGuard g(ix->device);
  int rc = ix->d_small.ensure(64);
  if (rc) return rc;
  bucket_lookup_kernel<<<1, 1>>>(ix->tab[table], index, (uint32_t*)ix->d_small.p);
  ix->launches++;
  uint32_t sl[2];
  CU(cudaMemcpy(sl, ix->d_small.p, 8, cudaMemcpyDeviceToHost));
  *n_out = sl[1];
  if (sl[1] == 0) return VC_NOT_FOUND;
  const uint32_t ncopy = std::min(cap, sl[1]);
  if (ids && ncopy) CU(cudaMemcpy(ids, ix->tab[table].ids + sl[0], (size_t)ncopy * 4, cudaMemcpyDeviceToHost));
  if (codes && ncopy) CU(cudaMemcpy(codes, ix->tab[table].codes + (size_t)sl[0] * ix->W, (size_t)ncopy * ix->W * 8, cudaMemcpyDeviceToHost));
  return VC_OK;
}

int vc_code_get(vc_index* ix, uint32_t id, void* code) {
  if (!ix || !code) return fail(VC_ERR_ARG, "null argument");
  if (id < ix->first_id || (id - ix->first_id) % ix->id_stride != 0 || (uint64_t)(id - ix->first_id) / ix->id_stride >= ix->n) return VC_NOT_FOUND;
  DeviceGuard g(ix->device);
  CU(cudaMemcpy(code, ix->d_codes + (uint64_t)((id - ix->first_id) / ix->id_stride) * ix->W, ix->W * 8, cudaMemcpyDeviceToHost));
  return VC_OK;
}

int vc_occupancy_bitmap_get(vc_index* ix, uint32_t table, uint32_t* words, uint64_t n_words) {
  if (!ix || !words) return fail(VC_ERR_ARG, "null argument");
  if (!ix->built || ix->m == 0) return fail(VC_ERR_STATE, "tables are not built");
  if (table >= ix->m) return fail(VC_ERR_ARG, "table %u out of range", table);
  const uint64_t need = 1ull << (ix->sbits >= 5 ? ix->sbits - 5 : 0);
  if (ix->sbits < 5 || n_words != need) return fail(VC_ERR_ARG, "bitmap of a %u-bit table has %llu words", ix->sbits, (unsigned long long)need);
  DeviceGuard g(ix->device);
  const TableDev& T = ix->tab[table];
  if (T.sparse) {
    CU(cudaMemcpy(words, T.bitmap, n_words * 4, cudaMemcpyDeviceToHost));
  } else {
    int rc = ix->d_small.ensure(n_words * 4);
    if (rc) return rc;
    dense_bitmap_kernel<<<grid_for(n_words, 256, ix->num_sms), 256>>>(T.row_ptr, n_words, (uint32_t*)ix->d_small.p);
    ix->launches++;
    CU(cudaMemcpy(words, ix->d_small.p, n_words * 4, cudaMemcpyDeviceToHost));
  }
  return VC_OK;
}

// ------------------------------------------------------------------------------------------------
// merge
// ------------------------------------------------------------------------------------------------
static uint32_t pow2_at_least(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

// lists [n_lists][nq][k] -> out [nq][k]; uses `scratch` (>= ceil(n_lists/fanin) * nq * k keys) for the tree levels
static int merge_lists(int64_t* launches, const uint64_t* d_lists, uint32_t n_lists, uint32_t nq, uint32_t k, uint32_t fanin,
                       uint64_t* d_scratch_a, uint64_t* d_scratch_b, uint64_t* d_out, cudaStream_t st, uint64_t list_stride = 0) {
  if (!list_stride) list_stride = (uint64_t)nq * k;
  uint32_t BUF = pow2_at_least(k + kMergeThreads);           // <= 4096 entries = 32 KB: no opt-in needed
  // a few short lists (the shards' top-k rows): all of them in the buffer at once, one selection pass (topk_compact_block)
  if ((uint64_t)std::min(n_lists, fanin) * k <= 2048) BUF = std::max(BUF, pow2_at_least(std::min(n_lists, fanin) * k));
  const size_t smem = (size_t)BUF * 8;
  const uint64_t* in = d_lists;
  uint32_t nl = n_lists;
  uint64_t* bufs[2] = {d_scratch_a, d_scratch_b};
  int flip = 0;
  while (nl > fanin) {
    const uint32_t groups = (nl + fanin - 1) / fanin;
    uint64_t* dst = bufs[flip];
    merge_topk_kernel<<<dim3(nq, groups), kMergeThreads, smem, st>>>(in, nl, fanin, nq, k, BUF, dst, list_stride);
    (*launches)++;
    in = dst; nl = groups; flip ^= 1; list_stride = (uint64_t)nq * k;
  }
  merge_topk_kernel<<<dim3(nq, 1), kMergeThreads, smem, st>>>(in, nl, nl, nq, k, BUF, d_out, list_stride);
  (*launches)++;
  CU(cudaGetLastError());
  return VC_OK;
}

int vc_merge_topk_dev(int device, const uint64_t* d_lists, uint32_t n_lists, uint32_t nq, uint32_t k, uint64_t* d_out, void* stream) {
  if (!d_lists || !d_out) return fail(VC_ERR_ARG, "null argument");
  if (k == 0 || k > VC_MAX_K) return fail(VC_ERR_ARG, "k must be in [1, %u]", VC_MAX_K);
  if (n_lists == 0 || nq == 0) return VC_OK;
  if (n_lists > 1024) return fail(VC_ERR_ARG, "at most 1024 lists per merge call");
  DeviceGuard g(device);
  if (!g.ok) return fail(VC_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  int64_t launches = 0;
  return merge_lists(&launches, d_lists, n_lists, nq, k, 1024, nullptr, nullptr, d_out, (cudaStream_t)stream);
}

int vc_merge_topk(int device, const uint64_t* lists, uint32_t n_lists, uint32_t nq, uint32_t k, uint64_t* out) {
  if (!lists || !out) return fail(VC_ERR_ARG, "null argument");
  DeviceGuard g(device);
  if (!g.ok) return fail(VC_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  uint64_t *d_in = nullptr, *d_out = nullptr;
  const size_t nin = (size_t)n_lists * nq * k, nout = (size_t)nq * k;
  CU(cudaMalloc(&d_in, std::max<size_t>(nin, 1) * 8));
  cudaError_t e = cudaMalloc(&d_out, std::max<size_t>(nout, 1) * 8);
  if (e != cudaSuccess) { cudaFree(d_in); return fail(VC_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); }
  cudaMemcpy(d_in, lists, nin * 8, cudaMemcpyHostToDevice);
  int rc = vc_merge_topk_dev(device, d_in, n_lists, nq, k, d_out, nullptr);
  if (rc == VC_OK) {
    e = cudaMemcpy(out, d_out, nout * 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(VC_ERR_CUDA, "cudaMemcpy: %s", cudaGetErrorString(e));
  }
  cudaFree(d_in); cudaFree(d_out);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// linear scan
// ------------------------------------------------------------------------------------------------
}  // extern "C"

template <int W, bool PF, int U4>
static int launch_bmih_verify(const BmihParams& p, int num_sms, cudaStream_t st, int* grid_io) {
  if (*grid_io == 0) {
    int occ = 1;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bmih_verify_kernel<W, PF, U4>, kBmihThreads, 0));
    *grid_io = std::max(1, occ) * num_sms;
  }
  bmih_verify_kernel<W, PF, U4><<<*grid_io, kBmihThreads, 0, st>>>(p);
  return VC_OK;
}

// tensor-core verify kernel: one persistent CTA per SM
template <int W, int QT>
static int launch_bmih_verify_tc(const BmihParams& p, int num_sms, cudaStream_t st) {
  using Cfg = TcCfg<W, QT>;
  CU(cudaFuncSetAttribute(bmih_verify_tc_kernel<W, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  bmih_verify_tc_kernel<W, QT><<<num_sms, Cfg::THREADS, Cfg::SMEM, st>>>(p);
  return VC_OK;
}
template <int W>
static int launch_bmih_verify_tc_any(const BmihParams& p, int num_sms, cudaStream_t st) {
  return launch_bmih_verify_tc<W, 128>(p, num_sms, st);
}
// the first version of the tensor-core kernel (tcverify_v1.cuh: A operand through shared memory, 4 expander warps, up to 256
// queries per tile): "scan.tc" / "mih.tc" = 2.  The faster of the two on large scan batches (profiles/tc_r02.md).
template <int W, int QT>
static int launch_bmih_verify_tc1(const BmihParams& p, int num_sms, cudaStream_t st) {
  using Cfg = Tc1Cfg<W, QT>;
  CU(cudaFuncSetAttribute(bmih_verify_tc1_kernel<W, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  bmih_verify_tc1_kernel<W, QT><<<num_sms, kTc1Threads, Cfg::SMEM, st>>>(p);
  return VC_OK;
}
template <int W> constexpr uint32_t tc1_max_qt() { return W == 4 ? 128u : 256u; }
template <int W>
static int launch_bmih_verify_tc1_any(const BmihParams& p, int num_sms, cudaStream_t st) {
  if constexpr (W == 4) return p.qt > 64 ? launch_bmih_verify_tc1<4, 128>(p, num_sms, st) : launch_bmih_verify_tc1<4, 64>(p, num_sms, st);
  else return p.qt > 64 ? launch_bmih_verify_tc1<W, 256>(p, num_sms, st) : launch_bmih_verify_tc1<W, 64>(p, num_sms, st);
}
template <int W> constexpr uint32_t tc_max_qt() { return 128u; }

static int search_linear_ring(vc_index* ix, const void* d_queries, uint32_t nq, uint32_t k, uint64_t* d_out_keys, cudaStream_t st);

// Brute-force scan of a large batch through the batched-MIH verify kernel (bmih.cuh): warp-granular work items of
// <= 16 K codes x <= 32 queries, launch-wide thresholds, per-query global candidate buffers, one settle pass.
// Falls back to the TMA-ring kernel if a candidate buffer overflows (heavy ties).
template <int W>
static int scan_batched(vc_index* ix, const void* d_queries, uint32_t nq, uint32_t k, uint64_t* d_out_keys, cudaStream_t st) {
  using Cfg = BmihCfg<W>;
  int rc;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const uint32_t cap = bmih_cap_for(k);
  const size_t o_gbuf = take((size_t)nq * cap * 8), o_taukey = take((size_t)nq * 8), o_hist = take((size_t)nq * Cfg::HB * 4),
               o_cnt = take((size_t)nq * 4), o_tau = take((size_t)nq * 4), o_flag = take((size_t)nq * 4), o_rad = take((size_t)nq * 4),
               o_probes = take((size_t)nq * 8), o_cands = take((size_t)nq * 8), o_actA = take((size_t)nq * 4), o_actB = take((size_t)nq * 4),
               o_xhist = take((size_t)nq * Cfg::HB * 4), o_ctr = take(128), o_tab = take(sizeof(TableDev));
  if ((rc = ix->b_state.ensure(off))) return rc;
  unsigned char* sb = (unsigned char*)ix->b_state.p;
  uint32_t* ctr = (uint32_t*)(sb + o_ctr);
  BmihParams p;
  memset(&p, 0, sizeof p);
  p.queries = (const uint32_t*)d_queries; p.nq = nq; p.k = k; p.m = 1; p.sbits = 0; p.max_radius = 0; p.cap = cap;
  p.spec_tau = kInfDist;
  p.scan_mode = 1; p.first_id = ix->first_id; p.id_stride = ix->id_stride;
  p.pf_tau = ix->scan_prefilter > 0 ? 0x7FFFFFFFu : bmih_pf_tau(W, 0, true);
  TableDev pseudo;
  memset(&pseudo, 0, sizeof pseudo);
  pseudo.codes = ix->d_codes;
  CU(cudaMemcpyAsync(sb + o_tab, &pseudo, sizeof pseudo, cudaMemcpyHostToDevice, st));
  p.tables = (const TableDev*)(sb + o_tab);
  p.n_items = ctr; p.item_cursor = ctr + 1; p.n_next = ctr + 2;
  p.bucket_codes = (unsigned long long*)(ctr + 8); p.pair_count = (unsigned long long*)(ctr + 10);
  p.tc_stats = (unsigned long long*)(ctr + 16);
  p.gbuf = (uint64_t*)(sb + o_gbuf); p.gcnt = (uint32_t*)(sb + o_cnt); p.gtaukey = (uint64_t*)(sb + o_taukey);
  p.gtau = (uint32_t*)(sb + o_tau); p.ghist = (uint32_t*)(sb + o_hist); p.gflag = (uint32_t*)(sb + o_flag);
  p.gradius = (uint32_t*)(sb + o_rad); p.gprobes = (unsigned long long*)(sb + o_probes); p.gcands = (unsigned long long*)(sb + o_cands);
  uint32_t* ident = (uint32_t*)(sb + o_actA);           // 0 .. nq-1: the query list of every item, and the settle list
  p.next_active = (uint32_t*)(sb + o_actB);
  uint32_t* xhist = (uint32_t*)(sb + o_xhist);
  p.qlist = ident;
  // >= scan.tc_min queries per code: the distance filter goes to the tensor cores (tcverify.cuh)
  // by size: 64-bit codes, >= scan.tc_min queries AND a shard of >= 2^29 codes.  The tensor-core kernel flags a ROW (one code against
  // all queries of the item) when any of them passes the largest threshold of the item, and takes the exact path for the whole
  // row: while the thresholds are still settling - the first tens of millions of codes of a scan, whatever its length - that is
  // far more expensive than the POPC kernel's per-pair hit path.  Measured: 1.13 - 1.16 x the POPC kernel on 1 B codes, 0.5 x on a
  // 125 M-code shard (profiles/tc_r02.md).
  const bool use_tc = ix->scan_tc > 0 || (ix->scan_tc < 0 && W == 1 && nq >= (uint32_t)ix->scan_tc_min && k <= 128 && ix->n >= (1ull << 29));
  const bool tc_v1 = ix->scan_tc == 2 || ix->scan_tc < 0;       // by size: the version that measured faster than the POPC kernel
  p.cpi = use_tc ? kTcCpi : 8 * Cfg::STEP;
  p.qt = use_tc ? (tc_v1 ? (nq > 64 ? tc1_max_qt<W>() : 64u) : tc_max_qt<W>()) : (uint32_t)kBmihQT;
  ix->last_scan_tc = use_tc ? 1 : 0;
  const uint64_t n = ix->n;
  const uint32_t nc = (uint32_t)((n + p.cpi - 1) / p.cpi), nqc = (nq + p.qt - 1) / p.qt;
  const uint64_t n_items64 = (uint64_t)nc * nqc;
  if (n_items64 >= 0xFFFFFFF0ull) return fail(VC_ERR_ARG, "batch too large for one scan call; split it");
  const uint32_t n_items = (uint32_t)n_items64;
  if ((rc = ix->b_items.ensure(std::max<size_t>(n_items, 1) * sizeof(BmihItem)))) return rc;
  p.items = (BmihItem*)ix->b_items.p;
  CU(cudaMemsetAsync(ctr, 0, 128, st));
  CU(cudaMemsetAsync(p.ghist, 0, (size_t)nq * Cfg::HB * 4, st));
  bmih_init_kernel<<<(nq + 255) / 256, 256, 0, st>>>(p, ident);
  scan_bootstrap_kernel<W><<<nq, 256, 0, st>>>(p, ix->d_codes, n);
  scan_items_kernel<<<grid_for(std::max<uint32_t>(n_items, 1), 256, ix->num_sms), 256, 0, st>>>(p.items, n, p.cpi, nq, nqc, n_items);
  CU(cudaMemcpyAsync(ctr, &n_items, 4, cudaMemcpyHostToDevice, st));
  const bool pf = ix->scan_prefilter < 0 ? (W <= 2) : ix->scan_prefilter != 0;
  int grid = 0;
  if (ix->profile) cudaEventRecord(ix->ev0, st);
  if (use_tc && ix->tc_trace) {
    if ((rc = ix->b_trace.ensure(3 * 256 * 4 * 8))) return rc;
    CU(cudaMemsetAsync(ix->b_trace.p, 0, 3 * 256 * 4 * 8, st));
    p.tc_trace = (long long*)ix->b_trace.p;
  }
  if (use_tc) rc = tc_v1 ? launch_bmih_verify_tc1_any<W>(p, ix->num_sms, st) : launch_bmih_verify_tc_any<W>(p, ix->num_sms, st);
  else rc = pf ? launch_bmih_verify<W, true, kBmihU4>(p, ix->num_sms, st, &grid) : launch_bmih_verify<W, false, kBmihU4>(p, ix->num_sms, st, &grid);
  if (rc) return rc;
  if (ix->profile) { cudaEventRecord(ix->ev1, st); ix->ev_valid = true; ix->lev_used = 0; }
  XchgDev nox;
  memset(&nox, 0, sizeof nox);
  bmih_settle_kernel<W><<<nq, 256, 0, st>>>(p, ident, nq, xhist, nox);
  bmih_finish_kernel<<<nq, 128, 0, st>>>(p, d_out_keys, nullptr, nox);
  ix->launches += 6;
  CU(cudaGetLastError());
  // any overflow?  (flags are written by the append path; one small read-back)
  std::vector<uint32_t> flags(nq);
  CU(cudaMemcpyAsync(flags.data(), p.gflag, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  unsigned long long tcs[3] = {0, 0, 0};
  CU(cudaMemcpyAsync(tcs, p.tc_stats, 24, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  ix->tc_units = (int64_t)tcs[0]; ix->tc_flagged = (int64_t)tcs[1]; ix->tc_hits = (int64_t)tcs[2];
  if (p.tc_trace) {
    std::vector<long long> tr(3 * 256 * 4);
    CU(cudaMemcpy(tr.data(), p.tc_trace, tr.size() * 8, cudaMemcpyDeviceToHost));
    const long long t0 = tr[0];
    fprintf(stderr, "tc trace (CTA 0)\n unit: mma: stage_free issued committed - | epilogue: t_full pass1 released hits_done | tile: expander starts, raw_full, a_empty, published\n");
    for (int i = 0; i < 256; ++i) {
      fprintf(stderr, "%3d:", i);
      for (int j = 0; j < 4; ++j) fprintf(stderr, " %8lld", tr[i * 4 + j] ? tr[i * 4 + j] - t0 : -1);
      fprintf(stderr, " |");
      for (int j = 0; j < 4; ++j) fprintf(stderr, " %8lld", tr[(256 + i) * 4 + j] ? tr[(256 + i) * 4 + j] - t0 : -1);
      fprintf(stderr, " |");
      for (int j = 0; j < 4; ++j) fprintf(stderr, " %8lld", tr[(512 + i) * 4 + j] ? tr[(512 + i) * 4 + j] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  bool overflow = false;
  for (uint32_t q = 0; q < nq && !overflow; ++q) overflow = (flags[q] & 1u) != 0;
  ix->last_scan_batched = overflow ? 0 : 1;
  if (overflow) return search_linear_ring(ix, d_queries, nq, k, d_out_keys, st);
  return VC_OK;
}

template <int W, bool PF>
static int launch_scan(vc_index* ix, const ScanParams& p, size_t smem, cudaStream_t st) {
  auto kern = scan_topk_kernel<W, PF>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_optin));
  if (ix->profile) cudaEventRecord(ix->ev0, st);
  kern<<<p.n_qtiles * p.n_slices, kScanCtaThreads, smem, st>>>(p);
  if (ix->profile) { cudaEventRecord(ix->ev1, st); ix->ev_valid = true; ix->lev_used = 0; }
  ix->launches++;
  CU(cudaGetLastError());
  return VC_OK;
}

template <int W, bool PF>
static int scan_occupancy(size_t smem, int* occ) {
  CU(cudaFuncSetAttribute(scan_topk_kernel<W, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, scan_topk_kernel<W, PF>, kScanCtaThreads, smem));
  return VC_OK;
}

extern "C" {

int vc_search_linear_dev(vc_index* ix, const void* d_queries, uint32_t nq, uint32_t k, uint64_t* d_out_keys, void* stream) {
  if (!ix || !d_queries || !d_out_keys) return fail(VC_ERR_ARG, "null argument");
  if (k == 0 || k > VC_MAX_K) return fail(VC_ERR_ARG, "k must be in [1, %u]", VC_MAX_K);
  if (nq == 0) return VC_OK;
  DeviceGuard g(ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  // a single query is HBM-bound: the TMA-ring kernel streams the shard once at ~0.91 of the copy peak; batches go through
  // the warp-granular verify kernel (HBM-bound up to ~4 queries, POPC-bound above)
  // (candidates appended per query ~ 15-20 k before the thresholds settle: automatic only while that fits the buffer)
  // (2 .. 7 queries only on shards of >= 4 GB: the batched path has ~0.5 ms of fixed work in front of its kernel - threshold
  // bootstrap, settle, one host read-back of the overflow flags - which a 0.3 ms scan of a small shard does not pay back)
  const bool batched = ix->scan_batched > 0 || (ix->scan_batched < 0 && k <= 128 &&
                       (nq >= 8 || (nq >= (uint32_t)ix->scan_batched_min && ix->n * ix->W * 8 >= (4ull << 30))));
  ix->last_scan_batched = 0;
  if (batched && k < (uint32_t)kBmihSort / 2 && ix->n > 0 && ix->n < 0xFFFFFFFFull) {
    if (ix->W == 1) return scan_batched<1>(ix, d_queries, nq, k, d_out_keys, st);
    if (ix->W == 2) return scan_batched<2>(ix, d_queries, nq, k, d_out_keys, st);
    return scan_batched<4>(ix, d_queries, nq, k, d_out_keys, st);
  }
  return search_linear_ring(ix, d_queries, nq, k, d_out_keys, st);
}

}  // extern "C"

static int search_linear_ring(vc_index* ix, const void* d_queries, uint32_t nq, uint32_t k, uint64_t* d_out_keys, cudaStream_t st) {
  const uint32_t W = ix->W;
  const int qstride = W == 1 ? ScanCfg<1>::QSTRIDE : W == 2 ? ScanCfg<2>::QSTRIDE : ScanCfg<4>::QSTRIDE;
  const int hb = W == 1 ? ScanCfg<1>::HB : W == 2 ? ScanCfg<2>::HB : ScanCfg<4>::HB;
  const uint32_t step = W == 1 ? ScanCfg<1>::STEP : W == 2 ? ScanCfg<2>::STEP : ScanCfg<4>::STEP;
  ScanParams p;
  p.codes = (const uint4*)ix->d_codes; p.n = ix->n; p.first_id = ix->first_id; p.id_stride = ix->id_stride;
  p.queries = (const uint32_t*)d_queries; p.nq = nq; p.k = k;
  p.BUF = pow2_at_least(k + kScanSub);
  p.compact_at = k + (p.BUF - k) / 2;
  // shared memory plan: [ring: stages x 16 KB][per-query state x QT].  Default: two CTAs per SM, the query
  // tile gets up to 64 KB, the ring whatever is left (2..8 stages).
  const size_t fixed = scan_smem_bytes(0, p.BUF, qstride, 0);
  const size_t per_q = scan_smem_bytes(1, p.BUF, qstride, 0) - fixed;
  const size_t cta_budget = (ix->smem_optin + 1024) / std::max<int64_t>(1, ix->scan_ctas_per_sm) - 2048;
  const size_t state_budget = ix->scan_smem_kb > 0 ? (size_t)ix->scan_smem_kb * 1024 : (size_t)64 * 1024;
  uint32_t qt = (uint32_t)std::max<size_t>(1, (std::max(state_budget, fixed + per_q) - fixed) / per_q);
  if (ix->scan_qt > 0) qt = (uint32_t)ix->scan_qt;
  qt = std::min(qt, nq);
  auto total_smem = [&](uint32_t q, uint32_t stg) { return scan_smem_bytes(q, p.BUF, qstride, stg); };
  uint32_t stages = 2;
  if (ix->scan_stages > 0) stages = (uint32_t)std::min<int64_t>(ix->scan_stages, kScanMaxStages);
  else while (stages < kScanMaxStages && total_smem(qt, stages + 1) <= cta_budget) ++stages;
  while (total_smem(qt, stages) > ix->smem_optin && stages > 2) --stages;
  while (total_smem(qt, stages) > ix->smem_optin && qt > 1) --qt;
  if (total_smem(qt, stages) > ix->smem_optin) return fail(VC_ERR_ARG, "k = %u needs more shared memory than the device has", k);
  // equal query tiles: 32 queries with room for 31 per CTA become 16 + 16, not 31 + 1
  {
    const uint32_t tiles = (nq + qt - 1) / qt;
    qt = (nq + tiles - 1) / tiles;
  }
  p.QT = qt; p.stages = stages;
  p.n_qtiles = (nq + qt - 1) / qt;
  const size_t smem = total_smem(qt, stages);
  // the half-POPC lower-bound test pays once the kernel is POPC-bound (a handful of queries per code)
  bool pf = ix->scan_prefilter < 0 ? (W <= 2 && nq >= 4) : ix->scan_prefilter != 0;
  int occ = 1, rc;
  if (W == 1) rc = pf ? scan_occupancy<1, true>(smem, &occ) : scan_occupancy<1, false>(smem, &occ);
  else if (W == 2) rc = pf ? scan_occupancy<2, true>(smem, &occ) : scan_occupancy<2, false>(smem, &occ);
  else rc = pf ? scan_occupancy<4, true>(smem, &occ) : scan_occupancy<4, false>(smem, &occ);
  if (rc) return rc;
  occ = std::max(occ, 1);
  const uint64_t capacity = (uint64_t)occ * ix->num_sms;
  const uint64_t waves = ix->scan_waves > 0 ? (uint64_t)ix->scan_waves : ((uint64_t)p.n_qtiles * 32 <= capacity ? 1 : 8);
  const uint64_t n_steps = std::max<uint64_t>(1, (ix->n + step - 1) / step);
  // one wave: never spill a few CTAs into a second, almost empty wave (round down); several waves: round up
  uint64_t slices = waves == 1 ? std::max<uint64_t>(1, capacity / p.n_qtiles)
                               : std::max<uint64_t>(1, (capacity * waves + p.n_qtiles - 1) / p.n_qtiles);
  slices = std::min(slices, n_steps);
  const uint64_t steps_per_slice = (n_steps + slices - 1) / slices;
  p.slice_codes = steps_per_slice * step;
  p.n_slices = (uint32_t)((n_steps + steps_per_slice - 1) / steps_per_slice);
  p.interleave = ix->scan_interleave ? 1u : 0u;
  ix->last_scan_grid = (int64_t)p.n_qtiles * p.n_slices; ix->last_scan_qt = qt; ix->last_scan_slices = p.n_slices;
  ix->last_scan_smem = (int64_t)smem; ix->last_scan_occ = occ; ix->last_scan_stages = stages;

  rc = ix->d_gstate.ensure((size_t)nq * (hb + 1) * 4);
  if (rc) return rc;
  p.gtau = (uint32_t*)ix->d_gstate.p;
  p.ghist = p.gtau + nq;            // nq is padded below to keep the histogram rows 16-byte aligned
  {
    const uint32_t nq_pad = (nq + 3) & ~3u;
    rc = ix->d_gstate.ensure(((size_t)nq_pad + (size_t)nq * hb) * 4);
    if (rc) return rc;
    p.gtau = (uint32_t*)ix->d_gstate.p;
    p.ghist = p.gtau + nq_pad;
    scan_init_kernel<<<grid_for((uint64_t)nq * hb, 256, ix->num_sms), 256, 0, st>>>(p.gtau, p.ghist, nq, (uint32_t)hb);
    ix->launches++;
  }
  const uint32_t fanin = (uint32_t)std::max<int64_t>(2, ix->merge_fanin);
  const size_t part_keys = (size_t)p.n_slices * nq * k;
  rc = ix->d_partial.ensure(part_keys * 8);
  if (rc) return rc;
  p.partial = (uint64_t*)ix->d_partial.p;
  uint64_t *sa = nullptr, *sb = nullptr;
  if (p.n_slices > fanin) {
    const size_t lvl1 = (size_t)((p.n_slices + fanin - 1) / fanin) * nq * k;
    const size_t lvl2 = (size_t)((lvl1 / ((size_t)nq * k) + fanin - 1) / fanin) * nq * k;
    rc = ix->d_partial2.ensure((lvl1 + lvl2) * 8);
    if (rc) return rc;
    sa = (uint64_t*)ix->d_partial2.p; sb = sa + lvl1;
  }
  if (W == 1) rc = pf ? launch_scan<1, true>(ix, p, smem, st) : launch_scan<1, false>(ix, p, smem, st);
  else if (W == 2) rc = pf ? launch_scan<2, true>(ix, p, smem, st) : launch_scan<2, false>(ix, p, smem, st);
  else rc = pf ? launch_scan<4, true>(ix, p, smem, st) : launch_scan<4, false>(ix, p, smem, st);
  if (rc) return rc;
  return merge_lists(&ix->launches, p.partial, p.n_slices, nq, k, fanin, sa, sb, d_out_keys, st);
}

// ------------------------------------------------------------------------------------------------
// MIH
// ------------------------------------------------------------------------------------------------
template <int W, bool APPROX>
static int launch_mih(vc_index* ix, const MihParams& p, cudaStream_t st) {
  auto kern = mih_search_kernel<W, APPROX>;
  const size_t smem = (size_t)p.BUFM * 8 + (size_t)kMihWarps * kMihWbuf * 8;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
  if (ix->profile) cudaEventRecord(ix->ev0, st);
  kern<<<p.nq, kMihThreads, smem, st>>>(p);
  if (ix->profile) { cudaEventRecord(ix->ev1, st); ix->ev_valid = true; ix->lev_used = 0; }
  ix->launches++;
  CU(cudaGetLastError());
  return VC_OK;
}

static int mih_per_query(vc_index* ix, const void* d_queries, uint32_t nq, uint32_t k, int approximate, int max_radius,
                         uint64_t* d_out_keys, vc_query_stats* d_stats, cudaStream_t st, const uint32_t* d_qsel = nullptr) {
  MihParams p;
  p.qsel = d_qsel;
  p.queries = (const uint32_t*)d_queries; p.nq = nq; p.k = k; p.m = ix->m; p.sbits = ix->sbits;
  p.BUFM = pow2_at_least(k + kMihWbuf);
  p.approximate = approximate; p.max_radius = max_radius;
  p.table_steps = ix->mih_table_steps != 0 ? 1 : 0;
  p.tables = ix->d_tab; p.out_keys = d_out_keys; p.stats = d_stats;
  const bool ap = approximate != 0 && max_radius < 0;
  if (ix->W == 1) return ap ? launch_mih<1, true>(ix, p, st) : launch_mih<1, false>(ix, p, st);
  if (ix->W == 2) return ap ? launch_mih<2, true>(ix, p, st) : launch_mih<2, false>(ix, p, st);
  return ap ? launch_mih<4, true>(ix, p, st) : launch_mih<4, false>(ix, p, st);
}

static uint64_t host_binom(uint32_t n, uint32_t r) {
  if (r > n) return 0;
  uint64_t c = 1;
  for (uint32_t i = 1; i <= r; ++i) c = c * (n - r + i) / i;
  return c;
}

// ---- peer exchange (xchg.cuh) -------------------------------------------------------------------------------
static uint64_t align16(uint64_t v) { return (v + 15) & ~(uint64_t)15; }
static bool xchg_fits(const vc_index* ix, uint64_t bytes) { return ix->x_open && ix->x_enabled && align16(bytes) <= ix->x_slot_bytes; }
// the next exchange: `bytes` per rank, published by `n_ctas` CTAs of the pushing kernel
static XchgDev xchg_begin(vc_index* ix, uint64_t bytes, uint32_t n_ctas) {
  XchgDev x;
  memset(&x, 0, sizeof x);
  for (uint32_t g = 0; g < ix->x_world; ++g) x.peer[g] = ix->x_peer[g];
  x.rank = ix->x_rank; x.world = ix->x_world;
  x.seq = ++ix->x_seq;
  x.stride = align16(bytes);
  x.slot_off = kXchgHeaderBytes + (uint64_t)(x.seq & 1u) * ix->x_world * ix->x_slot_bytes;
  ix->x_done_total += n_ctas;
  x.done = ix->x_ctr; x.done_target = ix->x_done_total; x.err = ix->x_ctr + 1;
  ++ix->last_xchg;
  return x;
}
static void xchg_push(vc_index* ix, const XchgDev& x, const void* d_src, uint64_t bytes, uint32_t grid, cudaStream_t st) {
  xchg_push_kernel<<<grid, 256, 0, st>>>(x, (const uint4*)d_src, align16(bytes) / 16);
  ix->launches++;
}
static uint32_t xchg_push_grid(const vc_index* ix, uint64_t bytes) {
  return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)ix->num_sms, (align16(bytes) / 16 + 255) / 256));
}
// Sum of the n_words u32 at d_words over all shards, in place: over peer memory when the windows are open and the payload
// fits a slot, else through the caller's all-reduce hook (NCCL).  Every shard makes the same choice (same sizes everywhere).
static bool xchg_allreduce_wanted(const vc_index* ix, uint64_t bytes) {
  if (!xchg_fits(ix, bytes)) return false;
  if (!ix->allreduce_fn) return true;                       // no hook: peer memory is the only way
  return ix->x_allreduce > 0 || (ix->x_allreduce < 0 && bytes * (ix->x_world - 1) <= (2ull << 20));
}
__global__ void emulate_sum_kernel(uint32_t* w, uint64_t n, uint32_t g) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] *= g;
}
static int shard_allreduce(vc_index* ix, uint32_t* d_words, uint64_t n_words, cudaStream_t st) {
  if (ix->x_emulate > 1) {
    emulate_sum_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(d_words, n_words, (uint32_t)ix->x_emulate);
    ix->launches++;
    return VC_OK;
  }
  if (xchg_allreduce_wanted(ix, n_words * 4)) {
    const uint32_t grid = xchg_push_grid(ix, n_words * 4);
    const XchgDev x = xchg_begin(ix, n_words * 4, grid);
    xchg_push(ix, x, d_words, n_words * 4, grid, st);
    xchg_sum_kernel<<<grid, 256, 0, st>>>(x, d_words, n_words);
    ix->launches++;
    return VC_OK;
  }
  if (!ix->allreduce_fn) return fail(VC_ERR_STATE, "id-sharded search: the payload (%llu bytes) does not fit the exchange window and no all-reduce hook is set", (unsigned long long)(n_words * 4));
  if (ix->allreduce_fn(ix->allreduce_user, d_words, n_words, (void*)st) != 0) return fail(VC_ERR_STATE, "all-reduce callback failed");
  return VC_OK;
}

// exclusive scan of up to 2048 * 2048 u32 in place, no allocation (sums: >= ceil(n / 2048) words)
static int scan_inplace_small(vc_index* ix, uint32_t* d, uint64_t n, uint32_t* sums, cudaStream_t st) {
  const uint64_t nb = (n + kScanTile - 1) / kScanTile;
  if (nb > (uint64_t)kScanTile) return fail(VC_ERR_ARG, "scan too large");
  if (nb <= 1) { scan_apply_kernel<<<1, kBuildThreads, 0, st>>>(d, n, nullptr, d); ix->launches++; return VC_OK; }
  scan_reduce_kernel<<<(unsigned)nb, kBuildThreads, 0, st>>>(d, n, sums);
  scan_apply_kernel<<<1, kBuildThreads, 0, st>>>(sums, nb, nullptr, sums);
  scan_apply_kernel<<<(unsigned)nb, kBuildThreads, 0, st>>>(d, n, sums, d);
  ix->launches += 3;
  return VC_OK;
}

// Bucket-stationary batched MIH (bmih.cuh).  Exact or fixed-radius search over dense tables.
constexpr uint32_t kSpecMinBatch = 1024;     // speculative thresholds: batches smaller than this neither use nor teach a guess (mih.speculate = 1)
template <int W>
static int mih_batched(vc_index* ix, const void* d_queries, uint32_t nq, uint32_t k, int max_radius,
                       uint64_t* d_out_keys, vc_query_stats* d_stats, cudaStream_t st) {
  using Cfg = BmihCfg<W>;
  const uint32_t m = ix->m, sbits = ix->sbits, n_buckets = m << sbits;
  int rc;
  // ---- workspace ---------------------------------------------------------------------------------------
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const uint32_t cap = ix->mih_cap > 0 ? (uint32_t)ix->mih_cap : bmih_cap_for(k);
  const size_t o_gbuf = take((size_t)nq * cap * 8), o_taukey = take((size_t)nq * 8), o_hist = take((size_t)nq * Cfg::HB * 4),
               o_cnt = take((size_t)nq * 4), o_tau = take((size_t)nq * 4), o_flag = take((size_t)nq * 4), o_rad = take((size_t)nq * 4),
               o_probes = take((size_t)nq * 8), o_cands = take((size_t)nq * 8), o_actA = take((size_t)nq * 4), o_actB = take((size_t)nq * 4),
               o_xhist = take((size_t)nq * Cfg::HB * 4), o_globkey = take((size_t)nq * 8),
               o_ctr = take(128);   // [0] n_items [1] item_cursor [2] n_next [3] any_overflow [4] n_likely [8..9] bucket_codes
  if ((rc = ix->b_state.ensure(off))) return rc;
  if ((rc = ix->b_buckets.ensure(((size_t)n_buckets * 2 + 2 + kScanTile) * 4 + 1024 + n_buckets))) return rc;
  unsigned char* sb = (unsigned char*)ix->b_state.p;
  uint32_t* ctr = (uint32_t*)(sb + o_ctr);            // [0] n_items  [1] item_cursor  [2] n_next  [3] any_overflow
  BmihParams p;
  memset(&p, 0, sizeof p);
  p.queries = (const uint32_t*)d_queries; p.nq = nq; p.k = k; p.m = m; p.sbits = sbits; p.radius = 0; p.max_radius = max_radius; p.cap = cap;
  p.spec_tau = kInfDist; p.spec_fail = ctr + 25;
  const bool spec_batch = max_radius < 0 && (ix->mih_speculate >= 2 || (ix->mih_speculate == 1 && nq >= kSpecMinBatch));
  if (max_radius < 0) {
    if (ix->mih_spec_force >= 0) p.spec_tau = (uint32_t)ix->mih_spec_force;
    else if (spec_batch && ix->spec_valid && ix->spec_k == k) p.spec_tau = ix->spec_tau;
  }
  ix->last_spec_tau = p.spec_tau == kInfDist ? -1 : (int64_t)p.spec_tau; ix->last_spec_fail = 0;
  p.tables = ix->d_tab; p.active = nullptr; p.n_active = 0;
  {
    // work-item length: about one average bucket, between 2 and 8 CTA steps
    const uint64_t avg = (ix->n >> sbits) + 1;
    const uint64_t steps = std::min<uint64_t>(8, std::max<uint64_t>(2, (avg + Cfg::STEP - 1) / Cfg::STEP));
    p.cpi = ix->mih_cpi_steps > 0 ? (uint32_t)ix->mih_cpi_steps * Cfg::STEP : (uint32_t)steps * Cfg::STEP;
    // 64-bit codes are loaded in pairs from an even position: an item that starts at an odd one is one code longer for the
    // kernel, and its hit queue numbers at most 64 warp steps (of 256 codes) per item
    if (W == 1) p.cpi -= 2;
  }
  // long buckets: 16 codes per thread per step (less loop overhead per pair); short ones: 8
  const bool wide = ix->mih_wide > 0;   // measured slower than the 8-codes-per-thread variant at 3 CTAs/SM; kept as a knob
  p.bcount = (uint32_t*)ix->b_buckets.p; p.boffs = p.bcount + n_buckets;
  uint32_t* scan_sums = p.boffs + n_buckets + 1 + 3;
  uint8_t* bflag = (uint8_t*)(scan_sums + kScanTile + 64);
  const int64_t r0_first = ix->mih_r0_first;
  p.qlist = nullptr; p.items = nullptr; p.n_items = ctr; p.item_cursor = ctr + 1; p.n_next = ctr + 2;
  p.bucket_codes = (unsigned long long*)(ctr + 8);
  p.pair_count = (unsigned long long*)(ctr + 10);
  p.exec_pairs = (unsigned long long*)(ctr + 12);
  p.scan_mode = 0; p.first_id = ix->first_id; p.id_stride = ix->id_stride;
  p.pf_tau = ix->mih_prefilter > 0 ? 0x7FFFFFFFu : bmih_pf_tau(W, sbits, false);
  p.gglobkey = ix->sharded() ? (uint64_t*)(sb + o_globkey) : nullptr;
  ix->last_xchg = 0;
  XchgDev nox;
  memset(&nox, 0, sizeof nox);
  p.boot_sample = (uint32_t)std::max<int64_t>(0, ix->mih_boot_sample);
  p.tc_stats = (unsigned long long*)(ctr + 16);
  const uint32_t popc_cpi = p.cpi;
  const bool tc_possible = ix->mih_tc != 0;
  p.qt = kBmihQT; p.cpi_alt = kTcCpi; p.qt_alt = 128; p.n_items_alt = tc_possible ? ctr + 5 : nullptr;
  ix->last_mih_tc_steps = 0; ix->last_mih_redo = 0;
  unsigned long long prev_codes = 0, prev_pairs = 0, prev_exec = 0;
  ix->step_codes.clear(); ix->step_pairs.clear(); ix->step_exec.clear();
  p.gbuf = (uint64_t*)(sb + o_gbuf); p.gcnt = (uint32_t*)(sb + o_cnt); p.gtaukey = (uint64_t*)(sb + o_taukey);
  p.gtau = (uint32_t*)(sb + o_tau); p.ghist = (uint32_t*)(sb + o_hist); p.gflag = (uint32_t*)(sb + o_flag);
  p.gradius = (uint32_t*)(sb + o_rad); p.gprobes = (unsigned long long*)(sb + o_probes); p.gcands = (unsigned long long*)(sb + o_cands);
  uint32_t* xhist = (uint32_t*)(sb + o_xhist);
  uint32_t* actA = (uint32_t*)(sb + o_actA);
  uint32_t* actB = (uint32_t*)(sb + o_actB);

  if (ix->max_bucket_len == 0) {
    CU(cudaMemsetAsync(ctr, 0, 4, st));
    bmih_maxlen_kernel<<<grid_for(n_buckets, 256, ix->num_sms), 256, 0, st>>>(ix->d_tab, m, sbits, ctr);
    uint32_t mx = 0;
    CU(cudaMemcpyAsync(&mx, ctr, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ix->max_bucket_len = std::max(mx, 1u);
    ix->launches++;
  }
  // ---- start: all queries active, thresholds bootstrapped from a sample of each query's own buckets -------
  if (ix->profile) {
    if (!ix->ev_s0) { CU(cudaEventCreate(&ix->ev_s0)); CU(cudaEventCreate(&ix->ev_s1)); }
    cudaEventRecord(ix->ev_s0, st);
  }
  CU(cudaMemsetAsync(ctr, 0, 128, st));
  CU(cudaMemsetAsync(p.ghist, 0, (size_t)nq * Cfg::HB * 4, st));      // histograms count this search's candidates only
  CU(cudaMemsetAsync(xhist, 0, (size_t)nq * Cfg::HB * 4, st));
  bmih_init_kernel<<<(nq + 255) / 256, 256, 0, st>>>(p, actA);
  bmih_bootstrap_kernel<W><<<(nq + 7) / 8, 256, 0, st>>>(p, ix->sharded() ? xhist : nullptr);
  ix->launches += 2;
  if (ix->sharded()) {
    if ((rc = shard_allreduce(ix, xhist, (uint64_t)nq * Cfg::HB, st))) return rc;
    bmih_boot_tau_kernel<W><<<(nq + 127) / 128, 128, 0, st>>>(p, xhist);
    ix->launches++;
  }
  uint32_t h_ctr[4] = {0, 0, 0, 0};
  uint32_t n_active = nq;
  uint32_t* cur = actA;
  uint32_t* nxt = actB;
  const bool pf_all = ix->mih_prefilter < 0 ? (W <= 2) : ix->mih_prefilter != 0;
  int verify_grid = 0;
  int levels = 0;
  int64_t items_total = 0;
  // thresholds straight from the bootstrap are loose - unless the speculative bound (the previous batch's largest k-th distance + 1)
  // caps them below what the lower-bound filter needs (bmih_pf_tau)
  bool loose_majority = !(p.spec_tau != kInfDist && p.spec_tau < p.pf_tau);
  // steps: (radius r, tables [t0, t1)).  A whole radius per step, or - when most queries are about to stop - one
  // table per step, so that the strict rule d_k <= m*r + t can end the search in the middle of a radius.
  uint32_t r = 0, t0 = 0;
  bool granular = ix->mih_table_steps > 0 && max_radius < 0;
  while (n_active > 0 && r <= sbits) {
    const uint32_t t1 = granular ? t0 + 1 : m;
    // radius 0 alone can only end a search whose k-th distance is below m; in the adaptive rhythm it is probed
    // together with radius 1 (one step less; its few, long work items overlap with radius 1's many)
    const uint32_t r_lo = r;
    if (r == 0 && !granular && ix->mih_table_steps < 0 && sbits >= 1 && max_radius != 0 && !ix->mih_split_r0) r = 1;
    p.radius = r; p.r_lo = r_lo; p.t_begin = t0; p.t_end = t1; p.active = cur; p.n_active = n_active; p.next_active = nxt;
    const bool timed = ix->profile && levels < 34;
    if (timed) {
      if (!ix->lev[2 * levels]) {
        CU(cudaEventCreate(&ix->lev[2 * levels])); CU(cudaEventCreate(&ix->lev[2 * levels + 1]));
        CU(cudaEventCreate(&ix->lev_x[2 * levels])); CU(cudaEventCreate(&ix->lev_x[2 * levels + 1]));
      }
      cudaEventRecord(ix->lev_x[2 * levels], st);
    }
    uint64_t per_table_probes = 0;
    for (uint32_t rr = r_lo; rr <= r; ++rr) per_table_probes += host_binom(sbits, rr);
    const uint64_t total_probes = (uint64_t)n_active * (t1 - t0) * per_table_probes;
    if (total_probes >= 0xFFFFFFF0ull) return fail(VC_ERR_ARG, "batch of %u queries needs %llu probes in one step; split the batch", nq, (unsigned long long)total_probes);
    if ((rc = ix->b_qlist.ensure(std::max<uint64_t>(total_probes, 1) * 4))) return rc;
    p.qlist = (uint32_t*)ix->b_qlist.p;
    const int pgrid = grid_for(total_probes, 256, ix->num_sms);
    CU(cudaMemsetAsync(p.bcount, 0, (size_t)n_buckets * 4, st));
    // a step that probes radius 0 together with higher radii: the queries' own buckets are verified first (bmih_items_kernel)
    const bool two_phase = r_lo == 0 && r > 0 && !tc_possible && pf_all &&
                           (r0_first > 0 || (r0_first < 0 && total_probes >= 2 * ((uint64_t)(t1 - t0) << sbits)));
    p.bflag = two_phase ? bflag : nullptr;
    if (two_phase) CU(cudaMemsetAsync(bflag, 0, n_buckets, st));
    bmih_probe_kernel<W><<<pgrid, 256, 0, st>>>(p, 0);
    CU(cudaMemcpyAsync(p.boffs, p.bcount, (size_t)n_buckets * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemsetAsync(p.boffs + n_buckets, 0, 4, st));
    if ((rc = scan_inplace_small(ix, p.boffs, (uint64_t)n_buckets + 1, scan_sums, st))) return rc;
    CU(cudaMemsetAsync(p.bcount, 0, (size_t)n_buckets * 4, st));
    bmih_probe_kernel<W><<<pgrid, 256, 0, st>>>(p, 1);
    CU(cudaMemsetAsync(ctr, 0, 12, st));
    CU(cudaMemsetAsync(ctr + 5, 0, 4, st));
    const int igrid = grid_for(n_buckets, 256, ix->num_sms);
    p.cpi = popc_cpi; p.qt = kBmihQT;
    // work items: normally ONE pass into a buffer sized by a bound that needs no device data (every probed bucket gives
    // <= ceil(longest bucket / cpi) x ceil(its queries / 32) items) - no counting pass, no host round trip; the counting
    // pass remains for the tensor-core choice and for degenerate tables whose bound would be huge
    const uint64_t nc_max = ((uint64_t)ix->max_bucket_len + p.cpi - 1) / p.cpi;
    const uint64_t item_bound = std::max<uint64_t>(1, nc_max) * (total_probes / kBmihQT + std::min<uint64_t>((uint64_t)(t1 - t0) << sbits, total_probes) + 1);
    const bool single_pass = !tc_possible && item_bound <= (64ull << 20);
    uint32_t h12[12] = {0};
    uint32_t n_items = 0;
    if (single_pass) {
      if ((rc = ix->b_items.ensure((size_t)item_bound * sizeof(BmihItem)))) return rc;
    } else {
    bmih_items_kernel<W><<<igrid, 256, 0, st>>>(p, 0);
    CU(cudaMemcpyAsync(h12, ctr, 48, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(h_ctr, h12, 16);
    n_items = h_ctr[0];
    }
    // which verify kernel?  pairs / codes = queries per code of this step: the POPC kernel pays for every pair, the
    // tensor-core kernel a flat >= 91 clocks per 128 codes for up to 64 queries
    bool step_tc = false;
    if (tc_possible) {
      unsigned long long cc, pp;
      memcpy(&cc, h12 + 8, 8); memcpy(&pp, h12 + 10, 8);
      const double codes_now = (double)(cc - prev_codes), pairs_now = (double)(pp - prev_pairs);
      step_tc = ix->mih_tc > 0 || (codes_now > 0 && pairs_now >= (double)ix->mih_tc_ratio * codes_now);
      if (step_tc) {
        p.cpi = kTcCpi;
        if (pairs_now >= 80.0 * codes_now && tc_max_qt<W>() > 128) {
          // long query lists: 256 (128) queries per item; the item count for that geometry needs its own counting pass
          p.qt = tc_max_qt<W>(); p.qt_alt = p.qt;
          CU(cudaMemsetAsync(ctr + 5, 0, 4, st));
          unsigned long long* keep_bc = p.bucket_codes; uint32_t* keep_n = p.n_items;
          p.bucket_codes = (unsigned long long*)(ctr + 6); p.n_items = ctr + 4;      // scratch: [4] is cleared before pass 1, [6..7] unused
          bmih_items_kernel<W><<<igrid, 256, 0, st>>>(p, 0);
          p.bucket_codes = keep_bc; p.n_items = keep_n; p.qt_alt = 128;
          CU(cudaMemcpyAsync(h12, ctr, 48, cudaMemcpyDeviceToHost, st));
          CU(cudaStreamSynchronize(st));
          ix->launches++;
        } else p.qt = 128;
        n_items = h12[5];
        ++ix->last_mih_tc_steps;
      }
    }
    if (!single_pass && (rc = ix->b_items.ensure(std::max<size_t>(n_items, 1) * sizeof(BmihItem)))) return rc;
    p.items = (BmihItem*)ix->b_items.p;
    CU(cudaMemsetAsync(ctr, 0, 12, st));
    CU(cudaMemsetAsync(ctr + 4, 0, 4, st));
    CU(cudaMemsetAsync(ctr + 14, 0, 4, st));
    p.count_in_write = single_pass ? 1u : 0u;
    if (two_phase) {
      bmih_items_kernel<W><<<igrid, 256, 0, st>>>(p, 1, 0);
      CU(cudaMemcpyAsync(ctr + 15, ctr, 4, cudaMemcpyDeviceToDevice, st));      // items of the radius-0 buckets: [0, ctr[15])
      bmih_items_kernel<W><<<igrid, 256, 0, st>>>(p, 1, 1);
      ix->launches++;
    } else {
      bmih_items_kernel<W><<<igrid, 256, 0, st>>>(p, 1);
    }
    if (timed) cudaEventRecord(ix->lev[2 * levels], st);
    // Which filter: the one-POPC-per-64-bit lower bound - unless most queries' thresholds are still too loose for it (fresh from
    // the bootstrap; counted by the decide kernel, bmih_pf_tau), or the step has fewer than two queries per probed bucket (it is
    // HBM-bound then, and the exact distance sends fewer codes down the hit path); 256-bit codes: the exact distance throughout
    // (the OR bound does not reject at the thresholds of config C5).
    const bool pf_auto = ix->mih_prefilter < 0;
    const bool pf = pf_all && !(pf_auto && (loose_majority || total_probes < 2 * ((uint64_t)(t1 - t0) << sbits)));
    if (two_phase && ix->mih_prefilter < 0 && !wide) {
      // the queries' own buckets hold their nearest candidates: verified first, on the exact distance (thresholds are fresh from
      // the bootstrap); after them every threshold is the query's radius-0 k-th distance, tight enough for the lower bound
      BmihParams p0 = p;
      p0.n_items = ctr + 15;
      rc = launch_bmih_verify<W, false, kBmihU4>(p0, ix->num_sms, st, &verify_grid);
      if (rc) return rc;
      CU(cudaMemcpyAsync(ctr + 1, ctr + 15, 4, cudaMemcpyDeviceToDevice, st));   // the cursor continues behind them
      rc = launch_bmih_verify<W, true, kBmihU4>(p, ix->num_sms, st, &verify_grid);
      ix->launches++;
    }
    else if (step_tc) rc = launch_bmih_verify_tc_any<W>(p, ix->num_sms, st);
    else if (wide) rc = pf ? launch_bmih_verify<W, true, 8>(p, ix->num_sms, st, &verify_grid) : launch_bmih_verify<W, false, 8>(p, ix->num_sms, st, &verify_grid);
    else rc = pf ? launch_bmih_verify<W, true, kBmihU4>(p, ix->num_sms, st, &verify_grid) : launch_bmih_verify<W, false, kBmihU4>(p, ix->num_sms, st, &verify_grid);
    if (rc) return rc;
    if (timed) cudaEventRecord(ix->lev[2 * levels + 1], st);
    // id-sharded search: the histograms are summed over the shards, so that every GPU filters and stops on the k-th
    // distance of the WHOLE database (and all ranks walk through the same steps).  Over peer memory the settle kernel
    // itself stores its rows into every shard's window and xchg_sum_kernel adds the G slots up; else the hook (NCCL).
    if (ix->sharded() && xchg_allreduce_wanted(ix, (uint64_t)nq * Cfg::HB * 4)) {
      const XchgDev x = xchg_begin(ix, (uint64_t)nq * Cfg::HB * 4, n_active);
      bmih_settle_kernel<W><<<n_active, 256, 0, st>>>(p, cur, n_active, xhist, x);
      xchg_sum_kernel<<<xchg_push_grid(ix, (uint64_t)nq * Cfg::HB * 4), 256, 0, st>>>(x, xhist, (uint64_t)nq * Cfg::HB);
      ix->launches++;
    } else {
      bmih_settle_kernel<W><<<n_active, 256, 0, st>>>(p, cur, n_active, xhist, nox);
      if (ix->sharded() && (rc = shard_allreduce(ix, xhist, (uint64_t)nq * Cfg::HB, st))) return rc;
    }
    bmih_decide_kernel<W><<<(n_active + 127) / 128, 128, 0, st>>>(p, cur, n_active, xhist, ctr + 3, ctr + 4, ctr + 14);
    if (timed) cudaEventRecord(ix->lev_x[2 * levels + 1], st);
    ix->launches += single_pass ? 6 : 7;
    CU(cudaGetLastError());
    uint32_t h5[16], x_err = 0;
    CU(cudaMemcpyAsync(h5, ctr, 64, cudaMemcpyDeviceToHost, st));
    if (ix->x_open) CU(cudaMemcpyAsync(&x_err, ix->x_ctr + 1, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (x_err) return fail(VC_ERR_STATE, "peer exchange timed out: a shard did not arrive");
    {
      unsigned long long cc, pp, xx;
      memcpy(&cc, h5 + 8, 8); memcpy(&pp, h5 + 10, 8); memcpy(&xx, h5 + 12, 8);
      ix->step_codes.push_back((int64_t)(cc - prev_codes)); ix->step_pairs.push_back((int64_t)(pp - prev_pairs));
      ix->step_exec.push_back((int64_t)(xx - prev_exec));
      prev_codes = cc; prev_pairs = pp; prev_exec = xx;
    }
    h_ctr[3] = h5[3];
    if (single_pass) n_items = h5[0];
    const uint32_t n_likely = h5[4];
    n_active = h5[2];
    loose_majority = (uint64_t)h5[14] * 2 >= n_active;
    std::swap(cur, nxt);
    ++levels;
    items_total += n_items;
    t0 = t1;
    if (t0 == m) {
      t0 = 0; ++r;
      // next radius: one table at a time if most of the remaining queries would stop inside it anyway
      granular = max_radius < 0 && (ix->mih_table_steps > 0 || (ix->mih_table_steps < 0 && (uint64_t)n_likely * 2 >= n_active));
    }
    if (ix->sharded() && granular && n_active > 0 && r <= sbits && ix->mih_global_key != 0) {
      // id-sharded search, next step one table of a radius: most queries' k-th distance now equals the distance bound of
      // what is left to find, so the k-th ID of the whole database decides what can still matter (bmih_idhist_kernel)
      const uint32_t lb_next = m * r + t0;
      const size_t words = (size_t)nq * (kIdBins + 1);
      if ((rc = ix->b_idh.ensure(words * 4))) return rc;
      uint32_t* idh = (uint32_t*)ix->b_idh.p;
      CU(cudaMemsetAsync(idh, 0, words * 4, st));
      bmih_idhist_kernel<<<(n_active * 32 + 255) / 256, 256, 0, st>>>(p, cur, n_active, lb_next, idh);
      if ((rc = shard_allreduce(ix, idh, (uint64_t)words, st))) return rc;
      bmih_idcut_kernel<<<(n_active + 127) / 128, 128, 0, st>>>(p, cur, n_active, lb_next, idh);
      ix->launches += 2;
    }
  }
  if (ix->profile) { ix->lev_used = std::min(levels, 34); ix->ev_valid = true; }
  if (max_radius < 0) {                 // the next batch's speculative threshold: this batch's largest final k-th distance
    bmih_maxtau_kernel<<<(nq + 255) / 256, 256, 0, st>>>(p, ctr + 24);
    ix->launches++;
  }
  // sharded call over peer memory (vc_search_sharded_dev): unless some query has to be redone below, the finish kernel
  // itself stores the result rows into every shard's window
  ix->x_pushed = false;
  if (ix->x_fuse_finish && !h_ctr[3] && xchg_fits(ix, (uint64_t)nq * k * 8)) {
    ix->x_finish = xchg_begin(ix, (uint64_t)nq * k * 8, nq);
    bmih_finish_kernel<<<nq, 128, 0, st>>>(p, d_out_keys, d_stats, ix->x_finish);
    ix->x_pushed = true;
  } else {
    bmih_finish_kernel<<<nq, 128, 0, st>>>(p, d_out_keys, d_stats, nox);
  }
  ix->launches++;
  if (ix->profile) cudaEventRecord(ix->ev_s1, st);
  if (h_ctr[3]) {
    // some candidate buffers overflowed (heavy ties): those queries - and only those - take the per-query kernel's exact answer
    if ((rc = ix->b_redo.ensure((size_t)nq * 4 + 256))) return rc;
    uint32_t* rl = (uint32_t*)ix->b_redo.p;
    CU(cudaMemsetAsync(ctr + 2, 0, 4, st));
    bmih_redo_list_kernel<<<(nq + 255) / 256, 256, 0, st>>>(p.gflag, nq, rl, ctr + 2);
    uint32_t n_redo = 0;
    CU(cudaMemcpyAsync(&n_redo, ctr + 2, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ix->launches++;
    ix->last_mih_redo = n_redo;
    if (n_redo && (rc = mih_per_query(ix, d_queries, n_redo, k, 0, max_radius, d_out_keys, d_stats, st, rl))) return rc;
    if (ix->profile) ix->lev_used = std::min(levels, 34);   // the per-query kernel reset it: the step timings stay those of the batched steps
  }
  CU(cudaGetLastError());
  unsigned long long h_bc = 0, tcs[3] = {0, 0, 0};
  uint32_t h_spec[2] = {0, 0};          // [0] largest final k-th distance, [1] queries the speculative threshold was too small for
  CU(cudaMemcpyAsync(&h_bc, p.bucket_codes, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(tcs, p.tc_stats, 24, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h_spec, ctr + 24, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (max_radius < 0) ix->last_spec_fail = h_spec[1];
  if (spec_batch) {
    // learn only from a batch in which the guess held for every query (a miss means the data moved: one batch without a guess
    // follows, which sees every k-th distance again); id-sharded: thresholds and misses come from the summed histograms, so
    // every shard learns the same value
    ix->spec_valid = h_spec[1] == 0 && h_spec[0] > 0 && h_spec[0] <= 64u * W;
    ix->spec_tau = h_spec[0]; ix->spec_k = k;      // no slack: one more unit of threshold triples the first step's hits (profiles/spec_r02.log)
  }
  ix->tc_units = (int64_t)tcs[0]; ix->tc_flagged = (int64_t)tcs[1]; ix->tc_hits = (int64_t)tcs[2];
  ix->last_mih_batched = 1; ix->last_mih_levels = levels; ix->last_mih_items = items_total; ix->last_mih_bucket_codes = (int64_t)h_bc;
  return VC_OK;
}

// Approximate mode (search_worker.cc:93-157) where buckets are long: the fixed-radius-0 search of the batched path answers every
// query that knows k * 20 distinct candidates after radius 0 (approx_radius0_kernel decides which, bmih.cuh); the others - none on
// a large uniform index - are answered by the per-query kernel, which widens the radius as the reference does.
template <int W>
static int mih_approx_batched(vc_index* ix, const void* d_queries, uint32_t nq, uint32_t k, uint64_t* d_out_keys, vc_query_stats* d_stats,
                              cudaStream_t st) {
  int rc;
  const size_t o_unique = 0, o_flag = (size_t)nq * 8, o_list = o_flag + (size_t)nq * 4, o_n = o_list + (size_t)nq * 4;
  if ((rc = ix->b_approx.ensure(o_n + 256))) return rc;
  unsigned char* base = (unsigned char*)ix->b_approx.p;
  unsigned long long* unique0 = (unsigned long long*)(base + o_unique);
  uint32_t* flag = (uint32_t*)(base + o_flag);
  uint32_t* list = (uint32_t*)(base + o_list);
  uint32_t* n_list = (uint32_t*)(base + o_n);
  approx_radius0_kernel<W><<<nq, 256, 0, st>>>((const uint32_t*)d_queries, ix->d_tab, ix->m, ix->sbits, nq,
                                               (unsigned long long)k * VC_APPROXIMATE_FACTOR, d_stats != nullptr ? 1 : 0, unique0, flag);
  ix->launches++;
  if ((rc = mih_batched<W>(ix, d_queries, nq, k, 0, d_out_keys, d_stats, st))) return rc;
  if (d_stats) { approx_stats_kernel<<<(nq + 255) / 256, 256, 0, st>>>(d_stats, unique0, flag, nq); ix->launches++; }
  CU(cudaMemsetAsync(n_list, 0, 4, st));
  bmih_redo_list_kernel<<<(nq + 255) / 256, 256, 0, st>>>(flag, nq, list, n_list);
  ix->launches++;
  uint32_t n_redo = 0;
  CU(cudaMemcpyAsync(&n_redo, n_list, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const int64_t tie_redo = ix->last_mih_redo;
  const int lev = ix->lev_used;
  if (n_redo && (rc = mih_per_query(ix, d_queries, n_redo, k, 1, -1, d_out_keys, d_stats, st, list))) return rc;
  ix->last_mih_redo = tie_redo + n_redo;
  if (ix->profile) ix->lev_used = lev;
  CU(cudaGetLastError());
  return VC_OK;
}

extern "C" {

int vc_search_mih_dev(vc_index* ix, const void* d_queries, uint32_t nq, uint32_t k, int approximate, int max_radius,
                      uint64_t* d_out_keys, vc_query_stats* d_stats, void* stream) {
  if (!ix || !d_queries || !d_out_keys) return fail(VC_ERR_ARG, "null argument");
  if (k == 0 || k > VC_MAX_K) return fail(VC_ERR_ARG, "k must be in [1, %u]", VC_MAX_K);
  if (ix->m == 0) return fail(VC_ERR_STATE, "index was created without tables (n_tables = 0)");
  if (!ix->built) return fail(VC_ERR_STATE, "tables are not built (call vc_index_build after adding codes)");
  if (nq == 0) return VC_OK;
  DeviceGuard g(ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  // Bucket-stationary batching pays when several queries share a bucket and buckets are long enough to
  // fill a CTA step; it needs dense tables and no distinct-candidate count (approximate mode).
  const bool legal = ix->sbits <= 16 && !(approximate != 0 && max_radius < 0) && k < (uint32_t)kBmihSort / 2;
  // id-sharded (an all-reduce hook is set): only the batched path calls the hook, so the choice must not depend on this
  // shard's size - shard sizes differ by one code when N % G != 0, and one rank issuing collectives the others do not is a hang
  const bool want = ix->mih_batched > 0 || (ix->mih_batched < 0 && (ix->sharded() || (ix->n >> ix->sbits) >= (uint64_t)ix->mih_min_bucket));
  ix->last_mih_batched = 0;
  if (legal && want) {
    if (ix->W == 1) return mih_batched<1>(ix, d_queries, nq, k, max_radius, d_out_keys, d_stats, st);
    if (ix->W == 2) return mih_batched<2>(ix, d_queries, nq, k, max_radius, d_out_keys, d_stats, st);
    return mih_batched<4>(ix, d_queries, nq, k, max_radius, d_out_keys, d_stats, st);
  }
  // approximate mode: batched while the queries' own buckets alone hold (twice) the k * 20 distinct candidates the mode stops at
  const bool approx_batched = approximate != 0 && max_radius < 0 && ix->sbits <= 16 && k < (uint32_t)kBmihSort / 2 && !ix->sharded() &&
                              (ix->mih_batched > 0 || (ix->mih_batched < 0 && (ix->n >> ix->sbits) * ix->m >= 2ull * k * VC_APPROXIMATE_FACTOR &&
                                                       (ix->n >> ix->sbits) >= (uint64_t)ix->mih_min_bucket));
  if (approx_batched) {
    if (ix->W == 1) return mih_approx_batched<1>(ix, d_queries, nq, k, d_out_keys, d_stats, st);
    if (ix->W == 2) return mih_approx_batched<2>(ix, d_queries, nq, k, d_out_keys, d_stats, st);
    return mih_approx_batched<4>(ix, d_queries, nq, k, d_out_keys, d_stats, st);
  }
  return mih_per_query(ix, d_queries, nq, k, approximate, max_radius, d_out_keys, d_stats, st);
}

// ------------------------------------------------------------------------------------------------
// host-buffer search entry points: H2D queries, search, unpack, D2H results
// ------------------------------------------------------------------------------------------------
static int search_host_chunk(vc_index* ix, bool mih, const void* queries, uint32_t nq, uint32_t k, int approximate, int max_radius,
                             uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_counts, vc_query_stats* stats);
// Host buffers of any length: the queries go through the device in chunks of `host.chunk` (32 768), so that the per-query
// workspace (candidate buffers, probe lists) stays bounded whatever the caller passes; one chunk = one batch of the batched paths.
static int search_host(vc_index* ix, bool mih, const void* queries, uint32_t nq, uint32_t k, int approximate, int max_radius,
                       uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_counts, vc_query_stats* stats) {
  if (!ix || (!queries && nq)) return fail(VC_ERR_ARG, "null argument");
  if (k == 0 || k > VC_MAX_K) return fail(VC_ERR_ARG, "k must be in [1, %u]", VC_MAX_K);
  const uint32_t chunk = (uint32_t)std::max<int64_t>(1, ix->host_chunk);
  for (uint32_t off = 0; off < nq; off += chunk) {
    const uint32_t n = std::min(chunk, nq - off);
    const int rc = search_host_chunk(ix, mih, (const unsigned char*)queries + (size_t)off * ix->W * 8, n, k, approximate, max_radius,
                                     out_ids ? out_ids + (size_t)off * k : nullptr, out_dists ? out_dists + (size_t)off * k : nullptr,
                                     out_counts ? out_counts + off : nullptr, stats ? stats + off : nullptr);
    if (rc) return rc;
  }
  return VC_OK;
}
static int search_host_chunk(vc_index* ix, bool mih, const void* queries, uint32_t nq, uint32_t k, int approximate, int max_radius,
                             uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_counts, vc_query_stats* stats) {
  if (nq == 0) return VC_OK;
  DeviceGuard g(ix->device);
  const size_t qbytes = (size_t)nq * ix->W * 8, rk = (size_t)nq * k;
  int rc;
  if ((rc = ix->d_q.ensure(qbytes)) || (rc = ix->h_q.ensure(qbytes)) || (rc = ix->d_keys.ensure(rk * 8)) ||
      (rc = ix->d_ids.ensure(rk * 4)) || (rc = ix->d_dists.ensure(rk * 4)) || (rc = ix->d_counts.ensure((size_t)nq * 4)) ||
      (rc = ix->h_ids.ensure(rk * 4)) || (rc = ix->h_dists.ensure(rk * 4)) || (rc = ix->h_counts.ensure((size_t)nq * 4)))
    return rc;
  if (stats && ((rc = ix->d_stats.ensure((size_t)nq * sizeof(vc_query_stats))) || (rc = ix->h_stats.ensure((size_t)nq * sizeof(vc_query_stats)))))
    return rc;
  cudaStream_t st = 0;
  memcpy(ix->h_q.p, queries, qbytes);
  CU(cudaMemcpyAsync(ix->d_q.p, ix->h_q.p, qbytes, cudaMemcpyHostToDevice, st));
  if (mih) rc = vc_search_mih_dev(ix, ix->d_q.p, nq, k, approximate, max_radius, (uint64_t*)ix->d_keys.p,
                                  stats ? (vc_query_stats*)ix->d_stats.p : nullptr, st);
  else rc = vc_search_linear_dev(ix, ix->d_q.p, nq, k, (uint64_t*)ix->d_keys.p, st);
  if (rc) return rc;
  unpack_keys_kernel<<<nq, 128, 0, st>>>((const uint64_t*)ix->d_keys.p, nq, k, (uint32_t*)ix->d_ids.p, (uint32_t*)ix->d_dists.p,
                                         (uint32_t*)ix->d_counts.p);
  ix->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ix->h_ids.p, ix->d_ids.p, rk * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(ix->h_dists.p, ix->d_dists.p, rk * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(ix->h_counts.p, ix->d_counts.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  if (stats) CU(cudaMemcpyAsync(ix->h_stats.p, ix->d_stats.p, (size_t)nq * sizeof(vc_query_stats), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (out_ids) memcpy(out_ids, ix->h_ids.p, rk * 4);
  if (out_dists) memcpy(out_dists, ix->h_dists.p, rk * 4);
  if (out_counts) memcpy(out_counts, ix->h_counts.p, (size_t)nq * 4);
  if (stats) memcpy(stats, ix->h_stats.p, (size_t)nq * sizeof(vc_query_stats));
  return VC_OK;
}

int vc_search_linear(vc_index* ix, const void* queries, uint32_t nq, uint32_t k, uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_counts) {
  return search_host(ix, false, queries, nq, k, 0, -1, out_ids, out_dists, out_counts, nullptr);
}

int vc_search_mih(vc_index* ix, const void* queries, uint32_t nq, uint32_t k, int approximate, int max_radius,
                  uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_counts, vc_query_stats* stats) {
  return search_host(ix, true, queries, nq, k, approximate, max_radius, out_ids, out_dists, out_counts, stats);
}

// ------------------------------------------------------------------------------------------------
// knobs
// ------------------------------------------------------------------------------------------------
int vc_index_set_allreduce(vc_index* ix, vc_allreduce_fn fn, void* user) {
  if (!ix) return fail(VC_ERR_ARG, "null argument");
  ix->allreduce_fn = fn;
  ix->allreduce_user = user;
  return VC_OK;
}

int vc_nccl_allreduce_hook(void* user, uint32_t* d_words, uint64_t n_words, void* stream) {
  const vc_nccl_hook* h = (const vc_nccl_hook*)user;
  if (!h || !h->nccl_allreduce || !h->comm) return fail(VC_ERR_ARG, "vc_nccl_allreduce_hook: no NCCL entry point or communicator");
  // ncclResult_t ncclAllReduce(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t); ncclUint32 = 3, ncclSum = 0
  typedef int (*nccl_allreduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  const int rc = ((nccl_allreduce_t)h->nccl_allreduce)(d_words, d_words, (size_t)n_words, 3, 0, h->comm, (cudaStream_t)stream);
  return rc == 0 ? VC_OK : fail(VC_ERR_STATE, "ncclAllReduce failed (ncclResult_t %d)", rc);
}

// ---- peer exchange ---------------------------------------------------------------------------------------------
int vc_xchg_create(vc_index* ix, uint32_t rank, uint32_t world, uint64_t slot_bytes, void* out_handle) {
  if (!ix) return fail(VC_ERR_ARG, "null argument");
  if (world < 2 || world > (uint32_t)kXchgMaxWorld || rank >= world) return fail(VC_ERR_ARG, "exchange of %u ranks (2 .. %d), rank %u", world, kXchgMaxWorld, rank);
  if (ix->x_window) return fail(VC_ERR_STATE, "the exchange window exists already");
  DeviceGuard g(ix->device);
  slot_bytes = align16(std::max<uint64_t>(slot_bytes, 4096));
  const uint64_t total = kXchgHeaderBytes + 2 * (uint64_t)world * slot_bytes;
  CU(cudaMalloc((void**)&ix->x_window, total));
  CU(cudaMemset(ix->x_window, 0, total));
  CU(cudaMalloc((void**)&ix->x_ctr, 64));
  CU(cudaMemset(ix->x_ctr, 0, 64));
  CU(cudaDeviceSynchronize());
  ix->x_rank = rank; ix->x_world = world; ix->x_slot_bytes = slot_bytes; ix->x_seq = 0; ix->x_done_total = 0; ix->x_open = false;
  if (out_handle) {
    cudaIpcMemHandle_t h;
    static_assert(sizeof(cudaIpcMemHandle_t) == VC_XCHG_HANDLE_BYTES, "IPC handle size");
    CU(cudaIpcGetMemHandle(&h, ix->x_window));
    memcpy(out_handle, &h, sizeof h);
  }
  return VC_OK;
}

int vc_xchg_local_window(vc_index* ix, void** window) {
  if (!ix || !window) return fail(VC_ERR_ARG, "null argument");
  if (!ix->x_window) return fail(VC_ERR_STATE, "no exchange window (vc_xchg_create)");
  *window = ix->x_window;
  return VC_OK;
}

int vc_xchg_open_ptrs(vc_index* ix, void* const* peer_windows) {
  if (!ix || !peer_windows) return fail(VC_ERR_ARG, "null argument");
  if (!ix->x_window) return fail(VC_ERR_STATE, "no exchange window (vc_xchg_create)");
  for (uint32_t g = 0; g < ix->x_world; ++g) {
    ix->x_peer[g] = g == ix->x_rank ? ix->x_window : (unsigned char*)peer_windows[g];
    ix->x_ipc[g] = false;
    if (!ix->x_peer[g]) return fail(VC_ERR_ARG, "window of rank %u is null", g);
  }
  ix->x_open = true;
  return VC_OK;
}

int vc_xchg_open(vc_index* ix, const void* handles) {
  if (!ix || !handles) return fail(VC_ERR_ARG, "null argument");
  if (!ix->x_window) return fail(VC_ERR_STATE, "no exchange window (vc_xchg_create)");
  DeviceGuard g(ix->device);
  for (uint32_t r = 0; r < ix->x_world; ++r) {
    if (r == ix->x_rank) { ix->x_peer[r] = ix->x_window; ix->x_ipc[r] = false; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * VC_XCHG_HANDLE_BYTES, sizeof h);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(VC_ERR_CUDA, "cudaIpcOpenMemHandle (window of rank %u): %s", r, cudaGetErrorString(e));
    ix->x_peer[r] = (unsigned char*)ptr; ix->x_ipc[r] = true;
  }
  ix->x_open = true;
  return VC_OK;
}

// One call = local search on this shard + exchange of the local top-k lists + merge: d_out_keys holds the answer over the
// WHOLE database on every rank.  Over peer memory when the windows are open (the batched MIH search pushes its rows from
// its last kernel, anything else through xchg_push_kernel); the caller falls back to its own all-gather otherwise.
int vc_search_sharded_dev(vc_index* ix, int mih, const void* d_queries, uint32_t nq, uint32_t k, int approximate, int max_radius,
                          uint64_t* d_out_keys, void* stream) {
  if (!ix || !d_queries || !d_out_keys) return fail(VC_ERR_ARG, "null argument");
  if (nq == 0) return VC_OK;
  const uint64_t bytes = (uint64_t)nq * k * 8;
  if (!xchg_fits(ix, bytes)) return fail(VC_ERR_STATE, "the exchange windows are not open or too small for %llu bytes per rank", (unsigned long long)bytes);
  DeviceGuard g(ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ix->x_local.ensure(bytes);
  if (rc) return rc;
  uint64_t* local = (uint64_t*)ix->x_local.p;
  ix->x_fuse_finish = mih != 0;
  ix->x_pushed = false;
  rc = mih ? vc_search_mih_dev(ix, d_queries, nq, k, approximate, max_radius, local, nullptr, stream)
           : vc_search_linear_dev(ix, d_queries, nq, k, local, stream);
  ix->x_fuse_finish = false;
  if (rc) return rc;
  XchgDev x;
  if (ix->x_pushed) {
    x = ix->x_finish;
  } else {
    const uint32_t grid = xchg_push_grid(ix, bytes);
    x = xchg_begin(ix, bytes, grid);
    xchg_push(ix, x, local, bytes, grid, st);
  }
  xchg_wait_kernel<<<1, 32, 0, st>>>(x);
  ix->launches++;
  const uint64_t* slots = (const uint64_t*)(ix->x_peer[ix->x_rank] + x.slot_off);
  rc = merge_lists(&ix->launches, slots, ix->x_world, nq, k, 1024, nullptr, nullptr, d_out_keys, st, x.stride / 8);
  if (rc) return rc;
  // a wait that timed out (a peer never arrived) must not pass as an answer
  uint32_t err = 0;
  CU(cudaMemcpyAsync(&err, ix->x_ctr + 1, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (err) return fail(VC_ERR_STATE, "peer exchange timed out: a shard did not arrive");
  return VC_OK;
}

int vc_index_set_param(vc_index* ix, const char* name, int64_t value) {
  if (!ix || !name) return fail(VC_ERR_ARG, "null argument");
  if (!strcmp(name, "scan.prefilter")) ix->scan_prefilter = value;
  else if (!strcmp(name, "scan.qt")) ix->scan_qt = value;
  else if (!strcmp(name, "scan.waves")) ix->scan_waves = value;
  else if (!strcmp(name, "scan.stages")) ix->scan_stages = value;
  else if (!strcmp(name, "scan.interleave")) ix->scan_interleave = value;
  else if (!strcmp(name, "scan.ctas_per_sm")) ix->scan_ctas_per_sm = value;
  else if (!strcmp(name, "scan.batched")) ix->scan_batched = value;
  else if (!strcmp(name, "scan.batched_min")) ix->scan_batched_min = value;
  else if (!strcmp(name, "scan.smem_kb")) ix->scan_smem_kb = value;
  else if (!strcmp(name, "merge.fanin")) ix->merge_fanin = value;
  else if (!strcmp(name, "id_stride")) {
    if (value < 1 || value > 65536) return fail(VC_ERR_ARG, "id_stride must be in [1, 65536]");
    if (ix->n) return fail(VC_ERR_STATE, "id_stride must be set before codes are added");
    ix->id_stride = (uint32_t)value;
  }
  else if (!strcmp(name, "tc.trace")) ix->tc_trace = value;
  else if (!strcmp(name, "scan.tc")) ix->scan_tc = value;
  else if (!strcmp(name, "scan.tc_min")) ix->scan_tc_min = value;
  else if (!strcmp(name, "mih.tc")) ix->mih_tc = value;
  else if (!strcmp(name, "mih.tc_ratio")) ix->mih_tc_ratio = value;
  else if (!strcmp(name, "mih.global_key")) ix->mih_global_key = value;
  else if (!strcmp(name, "xchg")) ix->x_enabled = value;
  else if (!strcmp(name, "xchg.allreduce")) ix->x_allreduce = value;
  else if (!strcmp(name, "xchg.emulate")) ix->x_emulate = value;
  else if (!strcmp(name, "mih.boot_sample")) ix->mih_boot_sample = value;
  else if (!strcmp(name, "mih.split_r0")) ix->mih_split_r0 = value;
  else if (!strcmp(name, "mih.speculate")) { ix->mih_speculate = value; ix->spec_valid = false; }
  else if (!strcmp(name, "mih.spec_tau")) ix->mih_spec_force = value;
  else if (!strcmp(name, "mih.r0_first")) ix->mih_r0_first = value;
  else if (!strcmp(name, "mih.cap")) {
    if (value < 0 || value > (1 << 20)) return fail(VC_ERR_ARG, "mih.cap must be in [0, 2^20]");
    ix->mih_cap = value;
  }
  else if (!strcmp(name, "mih.batched")) ix->mih_batched = value;
  else if (!strcmp(name, "mih.prefilter")) ix->mih_prefilter = value;
  else if (!strcmp(name, "mih.cpi_steps")) {
    if (value < 0 || value > 8) return fail(VC_ERR_ARG, "mih.cpi_steps must be in [0, 8]");     // the hit queue packs the warp step of the item into 6 bits
    ix->mih_cpi_steps = value;
  }
  else if (!strcmp(name, "mih.wide")) ix->mih_wide = value;
  else if (!strcmp(name, "mih.min_bucket")) ix->mih_min_bucket = value;
  else if (!strcmp(name, "host.chunk")) ix->host_chunk = std::max<int64_t>(1, value);
  else if (!strcmp(name, "mih.table_steps")) ix->mih_table_steps = value;
  else if (!strcmp(name, "profile")) {
    DeviceGuard g(ix->device);
    if (value && !ix->ev0) { CU(cudaEventCreate(&ix->ev0)); CU(cudaEventCreate(&ix->ev1)); }
    ix->profile = value; ix->ev_valid = false;
  }
  else return fail(VC_ERR_ARG, "unknown parameter '%s'", name);
  return VC_OK;
}

int vc_index_get_param(const vc_index* ix, const char* name, int64_t* value) {
  if (!ix || !name || !value) return fail(VC_ERR_ARG, "null argument");
  if (!strcmp(name, "scan.prefilter")) *value = ix->scan_prefilter;
  else if (!strcmp(name, "scan.qt")) *value = ix->scan_qt;
  else if (!strcmp(name, "scan.waves")) *value = ix->scan_waves;
  else if (!strcmp(name, "scan.smem_kb")) *value = ix->scan_smem_kb;
  else if (!strcmp(name, "merge.fanin")) *value = ix->merge_fanin;
  else if (!strcmp(name, "launches")) *value = ix->launches;
  else if (!strcmp(name, "scan.last_grid")) *value = ix->last_scan_grid;
  else if (!strcmp(name, "scan.last_qt")) *value = ix->last_scan_qt;
  else if (!strcmp(name, "scan.last_slices")) *value = ix->last_scan_slices;
  else if (!strcmp(name, "scan.last_smem")) *value = ix->last_scan_smem;
  else if (!strcmp(name, "scan.last_occ")) *value = ix->last_scan_occ;
  else if (!strcmp(name, "scan.last_batched")) *value = ix->last_scan_batched;
  else if (!strcmp(name, "scan.last_stages")) *value = ix->last_scan_stages;
  else if (!strcmp(name, "scan.stages")) *value = ix->scan_stages;
  else if (!strcmp(name, "num_sms")) *value = ix->num_sms;
  else if (!strcmp(name, "id_stride")) *value = ix->id_stride;
  else if (!strcmp(name, "scan.tc")) *value = ix->scan_tc;
  else if (!strcmp(name, "scan.last_tc")) *value = ix->last_scan_tc;
  else if (!strcmp(name, "mih.tc")) *value = ix->mih_tc;
  else if (!strcmp(name, "mih.tc_ratio")) *value = ix->mih_tc_ratio;
  else if (!strcmp(name, "mih.last_tc_steps")) *value = ix->last_mih_tc_steps;
  else if (!strcmp(name, "tc.last_units")) *value = ix->tc_units;
  else if (!strcmp(name, "tc.last_flagged")) *value = ix->tc_flagged;
  else if (!strcmp(name, "tc.last_hits")) *value = ix->tc_hits;
  else if (!strcmp(name, "mih.batched")) *value = ix->mih_batched;
  else if (!strcmp(name, "mih.last_batched")) *value = ix->last_mih_batched;
  else if (!strcmp(name, "mih.last_levels")) *value = ix->last_mih_levels;
  else if (!strcmp(name, "mih.last_redo")) *value = ix->last_mih_redo;
  else if (!strcmp(name, "mih.speculate")) *value = ix->mih_speculate;
  else if (!strcmp(name, "mih.last_spec_tau")) *value = ix->last_spec_tau;      // the speculative threshold of the last batched search (-1: none)
  else if (!strcmp(name, "mih.last_spec_fail")) *value = ix->last_spec_fail;    // queries it was too small for (redone by the per-query kernel)
  else if (!strcmp(name, "xchg")) *value = ix->x_enabled;
  else if (!strcmp(name, "xchg.open")) *value = ix->x_open ? 1 : 0;
  else if (!strcmp(name, "xchg.last")) *value = ix->last_xchg;
  else if (!strcmp(name, "mih.last_items")) *value = ix->last_mih_items;
  else if (!strcmp(name, "mih.last_bucket_codes")) *value = ix->last_mih_bucket_codes;
  else if (!strncmp(name, "mih.step_exec.", 14)) {
    const int i = atoi(name + 14);
    if (i < 0 || i >= (int)ix->step_exec.size()) return fail(VC_ERR_ARG, "step %d out of range", i);
    *value = ix->step_exec[i];
  }
  else if (!strncmp(name, "mih.step_codes.", 15) || !strncmp(name, "mih.step_pairs.", 15) || !strncmp(name, "mih.step_ns.", 12)) {
    // per-step accounting of the last batched search: mih.step_codes.<i>, mih.step_pairs.<i>, mih.step_ns.<i> (verify kernel time)
    const bool is_ns = name[9] == 'n';
    const int i = atoi(name + (is_ns ? 12 : 15));
    if (i < 0 || i >= (int)ix->step_codes.size()) return fail(VC_ERR_ARG, "step %d out of range", i);
    if (name[9] == 'c') *value = ix->step_codes[i];
    else if (name[9] == 'p') *value = ix->step_pairs[i];
    else {
      if (!ix->profile || i >= ix->lev_used) return fail(VC_ERR_STATE, "profiling is off");
      DeviceGuard g(ix->device);
      float ms = 0.f;
      CU(cudaEventSynchronize(ix->lev[2 * i + 1]));
      CU(cudaEventElapsedTime(&ms, ix->lev[2 * i], ix->lev[2 * i + 1]));
      *value = (int64_t)((double)ms * 1e6);
    }
  }
  else if (!strncmp(name, "mih.step_pre_ns.", 16) || !strncmp(name, "mih.step_post_ns.", 17)) {
    // the rest of a search step (profile = 1): begin of the step .. verify kernel (probes, counting sort, work items), and
    // verify kernel .. end of the step (settle, the cross-shard exchange, decide)
    const bool pre = name[10] == 'r';
    const int i = atoi(name + (pre ? 16 : 17));
    if (!ix->profile || i < 0 || i >= ix->lev_used) return fail(VC_ERR_STATE, "profiling is off or step out of range");
    DeviceGuard g(ix->device);
    float ms = 0.f;
    CU(cudaEventSynchronize(ix->lev_x[2 * i + 1]));
    if (pre) CU(cudaEventElapsedTime(&ms, ix->lev_x[2 * i], ix->lev[2 * i]));
    else CU(cudaEventElapsedTime(&ms, ix->lev[2 * i + 1], ix->lev_x[2 * i + 1]));
    *value = (int64_t)((double)ms * 1e6);
  }
  else if (!strcmp(name, "mih.search_ns")) {
    if (!ix->profile || !ix->ev_s0 || ix->lev_used <= 0) return fail(VC_ERR_STATE, "profiling is off or the last search was not batched");
    DeviceGuard g(ix->device);
    float ms = 0.f;
    CU(cudaEventSynchronize(ix->ev_s1));
    CU(cudaEventElapsedTime(&ms, ix->ev_s0, ix->ev_s1));
    *value = (int64_t)((double)ms * 1e6);
  }
  else if (!strcmp(name, "profile")) *value = ix->profile;
  else if (!strcmp(name, "last_kernel_ns")) {
    // device time of the dominant kernel (scan_topk / mih_search) of the last search; waits for it
    if (!ix->profile || !ix->ev_valid) return fail(VC_ERR_STATE, "profiling is off or no search has run");
    DeviceGuard g(ix->device);
    float ms = 0.f;
    if (ix->lev_used > 0) {               // batched MIH: sum of the verify kernels of all levels
      double tot = 0;
      for (int l = 0; l < ix->lev_used; ++l) {
        CU(cudaEventSynchronize(ix->lev[2 * l + 1]));
        CU(cudaEventElapsedTime(&ms, ix->lev[2 * l], ix->lev[2 * l + 1]));
        tot += ms;
      }
      *value = (int64_t)(tot * 1e6);
    } else {
      CU(cudaEventSynchronize(ix->ev1));
      CU(cudaEventElapsedTime(&ms, ix->ev0, ix->ev1));
      *value = (int64_t)((double)ms * 1e6);
    }
  }
  else return fail(VC_ERR_ARG, "unknown parameter '%s'", name);
  return VC_OK;
}

}  // extern "C"
