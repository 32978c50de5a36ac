/* verticut_gpu.h - C ABI of the B200-native VertiCut hot path (libverticut_gpu.so).
 *
 * This is the drop-in boundary for the MIH / linear-scan path of tu-dresden/verticut: plain
 * pointers and sizes, no C++ or torch types.  Each entry point names the reference interface it
 * replaces (paths relative to the reference tree).  The C++ mirror of the reference's own classes
 * (GpuTableProxy : BaseProxy, SearchWorker) lives in verticut_b200/host/ and calls only this ABI;
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success unless stated otherwise; a negative value is an error
 *    (VC_ERR_*), with text available from vc_last_error() (per host thread).
 *  - a code is `code_bits/8` raw bytes, exactly the record of the reference's code files
 *    (src/build_hash_tables.cc:40-45); the id of a code is `first_id` + its ordinal.
 *  - table t keys on bytes [t*s/8, (t+1)*s/8) of the code, little-endian UNSIGNED
 *    (Pilaf/image_tools.h:12-18 binaryToInt with its sign-extension bug removed).
 *  - distance = sum of popcounts of the XOR (Pilaf/image_tools.h:21-33).
 *  - search results are the k smallest packed words (dist << 32 | id)
 *    (src/search_worker.cc:12-13,254-256) in ASCENDING order - the canonical tie rule "by id".
 *    The reference emits descending distance (src/search_worker.cc:210-216); the C++ mirror
 *    reverses.  Unused result slots hold 0xFFFFFFFF / VC_EMPTY_KEY.
 *  - an index object may be used by one host thread at a time (the reference's proxies are
 *    single-threaded too: src/pilaf_proxy.h:19-21); batches provide the parallelism.
 *  - there is NO CPU fallback: without a CUDA device every call that computes fails with
 *    VC_ERR_CUDA.
 */
#ifndef VERTICUT_GPU_H
#define VERTICUT_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VC_ABI_VERSION 1

/* return codes.  0/1 mirror src/base_proxy.h:10-13 (PROXY_FOUND 0 / PROXY_NOT_FOUND 1 /
 * PROXY_PUT_DONE 0 / PROXY_PUT_FAIL 1). */
#define VC_OK 0
#define VC_NOT_FOUND 1
#define VC_ERR_ARG (-1)     /* bad argument (shape, k, table id, null pointer) */
#define VC_ERR_STATE (-2)   /* call not valid in this state (e.g. search before build) */
#define VC_ERR_CUDA (-3)    /* CUDA runtime error / no device */
#define VC_ERR_NOMEM (-4)   /* device or host allocation failed */

#define VC_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define VC_EMPTY_U32 0xFFFFFFFFu
#define VC_MAX_K 2048u
#define VC_APPROXIMATE_FACTOR 20u /* src/search_worker.h:14 */

typedef struct vc_index vc_index;

/* Per-query statistics; mirrors SearchWorker::get_stat (src/search_worker.cc:24-30). */
typedef struct vc_query_stats {
  uint32_t radius;           /* last substring radius searched (radius_, search_worker.cc:217) */
  uint32_t n_results;        /* entries returned (<= k) */
  uint64_t probes;           /* bucket lookups issued, all tables (n_sub_reads_, search_worker.cc:245) */
  uint64_t occupancy_tests;  /* occupancy-bitmap tests (n_local_reads_, search_worker.cc:239); s = 32 tables only */
  uint64_t candidates;       /* members of the probed buckets, duplicates included */
  uint64_t unique;           /* distinct candidates (knn_found_.size(), search_worker.cc:190); approximate mode only, else 0 */
} vc_query_stats;

typedef struct vc_index_info {
  uint64_t n_codes;
  uint32_t code_bits, n_tables, substring_bits, first_id;
  int32_t device;
  int32_t built;            /* tables are current */
  uint64_t device_bytes;    /* HBM held by the index, tables included */
} vc_index_info;

const char* vc_last_error(void);
int vc_abi_version(void);
/* number of CUDA devices visible, or VC_ERR_CUDA */
int vc_device_count(void);

/* ---- index life cycle ------------------------------------------------------------------
 * Replaces the KV servers behind BaseProxy::init/close (src/base_proxy.h:24-28): the tables live
 * in the HBM of `device`.  code_bits in {64,128,256}; n_tables = 0 builds no tables (linear scan
 * only), otherwise code_bits/n_tables must be 8, 16 or 32 (uint32 bucket index,
 * src/search_worker.cc:165-167).  first_id = global id of the first code of this shard. */
int vc_index_create(int device, uint32_t code_bits, uint32_t n_tables, uint32_t first_id, vc_index** out);
void vc_index_destroy(vc_index* ix);
int vc_index_get_info(const vc_index* ix, vc_index_info* info);

/* Appends n codes from host memory (raw records); ids continue from the last one.  Replaces the
 * fread loop of load_binarycode (src/build_hash_tables.cc:40-45,55,69).  Invalidates built tables. */
int vc_index_add(vc_index* ix, const void* codes, uint64_t n);
/* Appends n codes already in device memory of this index's device. */
int vc_index_add_device(vc_index* ix, const void* d_codes, uint64_t n);
/* Appends n synthetic codes generated on the device: 64-bit word w of code `id` is
 * vc_synth_word(seed, id, w) (same generator as oracle/verticut_oracle.c). */
int vc_index_add_synthetic(vc_index* ix, uint64_t n, uint64_t seed);
uint64_t vc_synth_word(uint64_t seed, uint64_t id, uint32_t word);

/* Builds all n_tables substring tables (CSR in HBM; occupancy bitmap + rank directory when s = 32).
 * Replaces the get/append/put loop of load_binarycode (src/build_hash_tables.cc:45-64) and
 * generate_bitmap (src/generate_bitmap.cc:99-125).  Bucket members end up in ascending id order. */
int vc_index_build(vc_index* ix);

/* Index persistence: the reference keeps its tables in external KV servers, so a search process finds them
 * built; here they live in the HBM of the process, so they can be written to / read from a file instead
 * (codes, all tables, occupancy bitmaps; little-endian, versioned header).  vc_index_load creates the index. */
int vc_index_save(vc_index* ix, const char* path);
int vc_index_load(int device, const char* path, vc_index** out);

/* ---- BaseProxy::get -------------------------------------------------------------------- */
/* get(HashIndex{table,index}, Image_List) (src/base_proxy.h:18, src/search_worker.cc:246): copies up
 * to `cap` members (ids and/or codes may be NULL) in stored order; *n = bucket size.
 * Returns VC_OK (found) / VC_NOT_FOUND (empty bucket). */
int vc_bucket_get(vc_index* ix, uint32_t table, uint32_t index, uint32_t* ids, void* codes, uint32_t cap, uint32_t* n);
/* get(ID{id}, BinaryCode) (src/linear_search.cc:45-46): VC_OK / VC_NOT_FOUND. */
int vc_code_get(vc_index* ix, uint32_t id, void* code);
/* Occupancy bitmap of a table in the reference's layout (bit i of word i/32, src/bitmap.cc:22-38;
 * one 2^s-bit map per table, src/generate_bitmap.cc:99-125).  n_words = 2^s / 32. */
int vc_occupancy_bitmap_get(vc_index* ix, uint32_t table, uint32_t* words, uint64_t n_words);

/* ---- search, host buffers (the reference-facing calls) --------------------------------------
 * queries: nq raw codes.  Outputs are [nq][k]; out_counts[q] <= k results are valid.  Any output
 * pointer may be NULL. */

/* Brute-force scan: search_K_nearest_neighbors(int k) of src/linear_search.cc:39-64 for a batch. */
int vc_search_linear(vc_index* ix, const void* queries, uint32_t nq, uint32_t k,
                     uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_counts);

/* MIH search: SearchWorker::find (src/search_worker.cc:65-89) for a batch.
 *  approximate = 0: exact k-NN (search_worker.cc:159-218) with the m-aware strict stop rule
 *                   d_k <= m*(r+1) - 1, by default also tested after every table t of a radius as
 *                   d_k <= m*r + t ("mih.table_steps"); results identical to vc_search_linear;
 *  approximate = 1: search_worker.cc:93-157 - stop at the first radius that has seen
 *                   >= 20*k distinct candidates, return the k best seen;
 *  max_radius >= 0: fixed-radius mode - search radii 0..max_radius, return the k best seen
 *                   (overrides both stop rules); max_radius < 0: use the stop rule. */
int vc_search_mih(vc_index* ix, const void* queries, uint32_t nq, uint32_t k, int approximate, int max_radius,
                  uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_counts, vc_query_stats* stats);

/* ---- search, device buffers (multi-GPU plumbing and zero-copy callers) ----------------------
 * d_queries / d_out_keys live on the index's device; d_out_keys is [nq][k] packed words
 * (VC_EMPTY_KEY padded), ascending.  `stream` is a cudaStream_t (NULL = default stream).  Neither call is
 * asynchronous: the MIH search walks its radius steps from the host and synchronises the stream between them
 * (it reads back a few counters per step); the linear scan only enqueues work for a single query, but a batch
 * goes through the batched kernels, which end with one stream synchronisation and a read-back of the per-query
 * overflow flags (a query whose candidate buffer overflowed - adversarial ties - is re-run by the ring kernel).
 * d_stats may be NULL.  In an id-sharded search (vc_index_set_allreduce / vc_xchg_open) every rank must make the
 * same calls in the same order; an error on one rank in the middle of a search (out of memory, a failing hook)
 * leaves the other ranks' exchanges unmatched and is fatal for the communicator / ends in the exchange's time-out. */
int vc_search_linear_dev(vc_index* ix, const void* d_queries, uint32_t nq, uint32_t k, uint64_t* d_out_keys, void* stream);
int vc_search_mih_dev(vc_index* ix, const void* d_queries, uint32_t nq, uint32_t k, int approximate, int max_radius,
                      uint64_t* d_out_keys, vc_query_stats* d_stats, void* stream);

/* Top-k merge of per-shard results; replaces mpi_coordinator::gather_vectors + the master's heap
 * (src/mpi_coordinator.cc:34-69, src/search_worker.cc:179-199) for id-disjoint shards.
 * d_lists is [n_lists][nq][k] (the layout an all-gather of per-GPU [nq][k] results produces);
 * d_out is [nq][k]. */
int vc_merge_topk_dev(int device, const uint64_t* d_lists, uint32_t n_lists, uint32_t nq, uint32_t k,
                      uint64_t* d_out, void* stream);
/* Host-buffer convenience form of the same merge (copies in, merges on `device`, copies out). */
int vc_merge_topk(int device, const uint64_t* lists, uint32_t n_lists, uint32_t nq, uint32_t k, uint64_t* out);

/* Id-sharded search over several GPUs (one process per GPU): `fn` must sum, in place and over all shards, the
 * n_words uint32 words at device address d_words (an all-reduce; enqueued on `stream` or ordered after it) and
 * return 0.  When set, the batched MIH search calls it once per search step on its per-query distance
 * histograms, which lets every shard filter and stop on the k-th distance of the whole database instead of its
 * own (fewer probes per shard; all shards then take the same steps, so the collective is matched), and, before a
 * table-granular step, on per-query id histograms that bound the k-th (distance, id) key of the whole database
 * ("mih.global_key", default 1).  Replaces the per-radius MPI_Gatherv / MPI_Bcast pair of
 * src/search_worker.cc:177,207.  fn = NULL (default): no exchange. */
typedef int (*vc_allreduce_fn)(void* user, uint32_t* d_words, uint64_t n_words, void* stream);
int vc_index_set_allreduce(vc_index* ix, vc_allreduce_fn fn, void* user);

/* A ready-made vc_allreduce_fn for hosts that hold an NCCL communicator: `user` points to a vc_nccl_hook with the address of
 * ncclAllReduce (the library does not link NCCL) and the communicator; the hook issues
 * ncclAllReduce(d_words, d_words, n_words, ncclUint32, ncclSum, comm, stream) and returns 0 on ncclSuccess.  With it the
 * per-step exchanges of the id-sharded search are enqueued from C, without a detour through the host language
 *   vc_index_set_allreduce(ix, vc_nccl_allreduce_hook, &hook);      (the hook struct must outlive its use) */
typedef struct vc_nccl_hook { void* nccl_allreduce; void* comm; } vc_nccl_hook;
int vc_nccl_allreduce_hook(void* user, uint32_t* d_words, uint64_t n_words, void* stream);

/* Peer exchange: the id-sharded search over the GPUs of ONE box without NCCL on the data path - every shard stores its
 * per-step histograms and its local top-k rows straight into the other GPUs' memory over NVLink / NVSwitch (CUDA IPC windows,
 * flags with system-scope release / acquire; verticut_b200/csrc/xchg.cuh).  Replaces src/mpi_coordinator.cc:26-69 and the
 * per-radius exchange of src/search_worker.cc:177,207 like the hook above does, and the all-gather + merge of the results too.
 *   vc_xchg_create      allocates this rank's window (two sets of `world` slots of slot_bytes each - the largest payload one
 *                       rank contributes to one exchange: max(nq * k * 8, nq * (64 * W + 32) * 4, nq * 257 * 4) for the batches to
 *                       come; larger payloads fall back to the hook) and returns its CUDA IPC handle (VC_XCHG_HANDLE_BYTES
 *                       bytes) for the other processes; out_handle may be NULL when all shards live in one process
 *   vc_xchg_open        opens the windows of all ranks from their handles (world x VC_XCHG_HANDLE_BYTES bytes, rank order; the
 *                       entry of this rank is ignored); vc_xchg_open_ptrs does the same from device pointers (one process)
 *   vc_search_sharded_dev   local search (mih = 1: vc_search_mih_dev, 0: vc_search_linear_dev) + exchange of the local
 *                       top-k + merge kernel: d_out_keys [nq][k] is the answer over the whole database, identical on every
 *                       rank.  All ranks must call it with the same arguments, in the same order, on streams that make progress
 *                       concurrently (a rank waits for its peers inside the call; a peer that never arrives ends the wait
 *                       with VC_ERR_STATE after 20 s).  Knob "xchg" = 0 turns the peer path off (hook / NCCL only). */
#define VC_XCHG_HANDLE_BYTES 64
int vc_xchg_create(vc_index* ix, uint32_t rank, uint32_t world, uint64_t slot_bytes, void* out_handle);
int vc_xchg_local_window(vc_index* ix, void** window);
int vc_xchg_open(vc_index* ix, const void* handles);
int vc_xchg_open_ptrs(vc_index* ix, void* const* peer_windows);
int vc_search_sharded_dev(vc_index* ix, int mih, const void* d_queries, uint32_t nq, uint32_t k, int approximate, int max_radius,
                          uint64_t* d_out_keys, void* stream);

/* Knobs and counters (integers).  Unknown names return VC_ERR_ARG.
 *   "id_stride"   id of the j-th code added = first_id + j * id_stride (default 1; G for one of G interleaved shards;
 *                 must be set before codes are added; stored by vc_index_save)
 *   "scan.*", "mih.*"   kernel selection and tuning, e.g. "scan.batched", "scan.prefilter", "mih.batched",
 *                 "mih.table_steps", "mih.global_key"; "scan.tc" = -1 (default): 64-bit scans of >= "scan.tc_min" (256) queries
 *                 over >= 2^29 codes use the tensor-core kernel, 0 never, 1 / 2 force a version; "mih.tc" = 1 routes the MIH
 *                 distance filter through it (off: measured slower)
 *   "mih.speculate" = 1 (default; 0 off, 2 also for batches below 1024 queries): the batched exact search caps every query's
 *                 starting threshold by the largest k-th distance of the previous (large) batch on this index; a query the guess
 *                 is too small for is detected and redone exactly
 *                 ("mih.spec_tau" >= 0 forces a guess; "mih.last_spec_tau" / "mih.last_spec_fail" report it; id-sharded: the same
 *                 setting on every rank - the guess and its misses are then derived from the summed histograms and agree)
 *   "host.chunk"  vc_search_linear / vc_search_mih pass their queries through the device in batches of this many (32 768)
 *   "xchg", "xchg.allreduce", "xchg.emulate"   the peer-memory exchange of an id-sharded search
 *   "profile", "last_kernel_ns", "launches", "mih.step_*", "tc.last_*"   read-back of timings and statistics */
int vc_index_set_param(vc_index* ix, const char* name, int64_t value);
int vc_index_get_param(const vc_index* ix, const char* name, int64_t* value);

#ifdef __cplusplus
}
#endif
#endif /* VERTICUT_GPU_H */
