// TEST INFRASTRUCTURE ONLY (oracle/_ref).  Never linked into, imported by or
// executed from the product path; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library built
// from this file.
//
// What it is: a thin driver around the reference's OWN hot-path sources,
// compiled UNMODIFIED from /root/reference by oracle/Makefile:
//   src/search_worker.cc  src/mpi_coordinator.cc  src/bitmap.cc  src/timer.cc
//   src/args_config.cc    src/build_hash_tables.cc (main renamed by -D)
//   src/integrity_check.cc (main renamed by -D)   src/linear_search.cc (main renamed by -D)
//   Pilaf/image_tools.h   src/pilaf_proxy.h  src/base_proxy.h
// The pieces of the reference that cannot exist in this container are replaced
// below the hot path by the shims in oracle/shim/: MPI ranks are threads
// (mpi.h), the Pilaf RDMA store is an in-memory byte KV (store-client.h), the
// protobuf messages are hand-written (image_search.pb.h).  No arithmetic of
// the path (binaryToInt, compute_hamming_dist, enumerate_entry, the heaps, the
// stop rule, the build loop) is restated here - it all runs from the
// reference's files.
//
// Deviations from the shipped reference, each required to get a usable oracle
// (SURVEY.md section 8(c)):
//   1. compiled with -funsigned-char, which turns binaryToInt
//      (Pilaf/image_tools.h:12-18) into the unsigned little-endian load it was
//      meant to be for substrings shorter than 4 bytes; identical for s = 32.
//   2. the "main table" id -> code (needed by linear_search.cc:45-46, no longer
//      written by build_hash_tables.cc) is filled by ref_put_main_table().
//   3. the stop rule keeps the reference's hard-coded `radius * 4`
//      (search_worker.cc:204); it is only distance-exact for m = 4 tables.
#include <assert.h>
#include <fcntl.h>
#include <getopt.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/wait.h>

#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "search_worker.h"   // the reference's own header (pulls pilaf_proxy.h, mpi_coordinator.h, shims)
#include "args_config.h"     // the reference's globals: image_total, binary_bits, ...
#include "image_tools.h"     // the reference's binaryToInt / compute_hamming_dist (Pilaf/image_tools.h)

// ---------------------------------------------------------------------------
// shim: MPI over threads
// ---------------------------------------------------------------------------
namespace {
const int kMaxRanks = 64;
int g_world = 1;
thread_local int t_rank = 0;
pthread_barrier_t g_bar;
bool g_bar_live = false;
const void* g_slot[kMaxRanks];

size_t type_size(MPI_Datatype t) { return t == MPI_LONG_LONG ? 8 : 4; }
void rendezvous() { if (g_world > 1) pthread_barrier_wait(&g_bar); }
}  // namespace

extern "C" {
void vc_shim_mpi_set_world(int size) {
  if (size < 1 || size > kMaxRanks) { fprintf(stderr, "oracle shim: bad world size %d\n", size); abort(); }
  if (g_bar_live) { pthread_barrier_destroy(&g_bar); g_bar_live = false; }
  g_world = size;
  if (size > 1) { pthread_barrier_init(&g_bar, 0, size); g_bar_live = true; }
}
void vc_shim_mpi_set_rank(int rank) { t_rank = rank; }

int MPI_Init(int*, char***) { return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int code) { fprintf(stderr, "oracle shim: MPI_Abort(%d)\n", code); abort(); return 0; }
int MPI_Barrier(MPI_Comm) { rendezvous(); return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = g_world; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = t_rank; return MPI_SUCCESS; }

int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm) {
  if (t_rank == root) g_slot[0] = buf;
  rendezvous();
  if (t_rank != root) memcpy(buf, g_slot[0], count * type_size(type));
  rendezvous();
  return MPI_SUCCESS;
}

int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
               MPI_Datatype recvtype, int root, MPI_Comm) {
  g_slot[t_rank] = sendbuf;
  rendezvous();
  if (t_rank == root)
    for (int r = 0; r < g_world; ++r)
      memcpy((char*)recvbuf + (size_t)r * recvcount * type_size(recvtype), g_slot[r], sendcount * type_size(sendtype));
  rendezvous();
  return MPI_SUCCESS;
}

int MPI_Gatherv(const void* sendbuf, int, MPI_Datatype, void* recvbuf, const int* recvcounts, const int* displs,
                MPI_Datatype recvtype, int root, MPI_Comm) {
  g_slot[t_rank] = sendbuf;
  rendezvous();
  if (t_rank == root)
    for (int r = 0; r < g_world; ++r)   // rank order, as MPI_Gatherv delivers it
      memcpy((char*)recvbuf + (size_t)displs[r] * type_size(recvtype), g_slot[r], (size_t)recvcounts[r] * type_size(recvtype));
  rendezvous();
  return MPI_SUCCESS;
}

int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype, MPI_Op, int root, MPI_Comm) {
  g_slot[t_rank] = sendbuf;   // only MPI_BOR over MPI_INT is used by the reference (mpi_coordinator.cc:17-19)
  rendezvous();
  if (t_rank == root) {
    int* out = (int*)recvbuf;
    for (int i = 0; i < count; ++i) out[i] = 0;
    for (int r = 0; r < g_world; ++r)
      for (int i = 0; i < count; ++i) out[i] |= ((const int*)g_slot[r])[i];
  }
  rendezvous();
  return MPI_SUCCESS;
}
}  // extern "C"

// ---------------------------------------------------------------------------
// shim: the Pilaf store as an in-memory KV of serialised bytes
// ---------------------------------------------------------------------------
namespace {
std::unordered_map<std::string, std::string> g_kv;
std::shared_mutex g_kv_mu;
}  // namespace

extern "C" void vc_shim_kv_clear(void) { std::unique_lock<std::shared_mutex> l(g_kv_mu); g_kv.clear(); }
extern "C" uint64_t vc_shim_kv_size(void) { std::shared_lock<std::shared_mutex> l(g_kv_mu); return g_kv.size(); }

int Client::put_with_size(const char* key, const char* value, size_t key_len, size_t val_len) {
  std::unique_lock<std::shared_mutex> l(g_kv_mu);
  g_kv[std::string(key, key_len)].assign(value, val_len);
  return 0;
}
int Client::get_with_size(const char* key, char* value, size_t key_len, size_t& val_len) {
  std::shared_lock<std::shared_mutex> l(g_kv_mu);
  auto it = g_kv.find(std::string(key, key_len));
  if (it == g_kv.end()) return POST_GET_MISSING;
  val_len = it->second.size();
  memcpy(value, it->second.data(), val_len);
  return POST_GET_FOUND;
}

// ---------------------------------------------------------------------------
// entry points of the reference translation units
// ---------------------------------------------------------------------------
int ref_build_tables_main(int argc, char* argv[]);      // src/build_hash_tables.cc:76 (main, renamed with -Dmain=)
int ref_integrity_check_main(int argc, char* argv[]);   // src/integrity_check.cc main, renamed
void search_K_nearest_neighbors(int k);                 // src/linear_search.cc:39-64
extern std::string search_code;                         // src/linear_search.cc:25
extern BaseProxy<google::protobuf::Message, google::protobuf::Message>* proxy_clt;  // src/linear_search.cc:26

namespace {
typedef BaseProxy<google::protobuf::Message, google::protobuf::Message> Proxy;
const char* kCfg = "/nonexistent/vc_oracle_pilaf.cnf";   // ConfigReader shim treats a missing file as "no servers"

Proxy* new_proxy() {
  Proxy* p = new PilafProxy<google::protobuf::Message, google::protobuf::Message>;   // the reference's class
  if (p->init(kCfg) != 0) { fprintf(stderr, "oracle: proxy init failed\n"); abort(); }
  return p;
}

int run_main_per_rank(int (*fn)(int, char**), const char* file, int bits, int tables) {
  // The reference runs `mpirun -n <tables>`; ranks do not talk to each other in
  // these two programs and share the args_config globals (image_total is the id
  // counter, build_hash_tables.cc:55,69), so the ranks are run one after another.
  vc_shim_mpi_set_world(tables);
  char bbuf[32], nbuf[32];
  snprintf(bbuf, sizeof bbuf, "%d", bits);
  snprintf(nbuf, sizeof nbuf, "%d", tables);
  for (int r = 0; r < tables; ++r) {
    vc_shim_mpi_set_rank(r);
    const char* argv_c[] = {"ref", "-s", "pilaf", "-c", kCfg, "-b", bbuf, "-n", nbuf, "-f", file, 0};
    char* argv[12];
    for (int i = 0; i < 12; ++i) argv[i] = (char*)argv_c[i];
    optind = 0;   // full getopt re-initialisation between calls (glibc)
    int rc = fn(11, argv);
    if (rc != 0) return rc;
  }
  vc_shim_mpi_set_rank(0);
  return 0;
}

// Runs fn() with stdout redirected into a temporary file; returns its content.
template <class F>
std::string capture_stdout(F fn) {
  fflush(stdout);
  char path[] = "/tmp/vc_oracle_out_XXXXXX";
  int fd = mkstemp(path);
  if (fd < 0) { perror("mkstemp"); abort(); }
  unlink(path);
  int saved = dup(1);
  dup2(fd, 1);
  fn();
  fflush(stdout);
  dup2(saved, 1);
  close(saved);
  std::string out;
  off_t len = lseek(fd, 0, SEEK_END);
  lseek(fd, 0, SEEK_SET);
  out.resize((size_t)len);
  size_t got = 0;
  while (got < (size_t)len) {
    ssize_t r = read(fd, &out[got], (size_t)len - got);
    if (r <= 0) break;
    got += (size_t)r;
  }
  close(fd);
  return out;
}
}  // namespace


// ---------------------------------------------------------------------------
// MemProxy: a second BaseProxy backend for the reference code, used where the
// byte-KV round trip above would make a CPU baseline needlessly slow.  Same
// interface (src/base_proxy.h:15-29); values are kept as parsed messages and
// codes as a flat caller-owned array.  get(ID) and get(HashIndex) are what the
// reference's linear scan (linear_search.cc:46) and enumerate_entry
// (search_worker.cc:246) call.
// ---------------------------------------------------------------------------
namespace {
class MemProxy : public Proxy {
 public:
  const uint8_t* codes = 0;      // main table, borrowed
  uint64_t n = 0;
  int nbytes = 0;
  uint32_t first_id = 0;
  std::vector<std::unordered_map<uint32_t, Image_List> > tables;

  int get(const google::protobuf::Message& key, google::protobuf::Message& value) {
    if (const HashIndex* h = dynamic_cast<const HashIndex*>(&key)) {
      if (h->table_id() >= tables.size()) return PROXY_NOT_FOUND;
      auto it = tables[h->table_id()].find(h->index());
      if (it == tables[h->table_id()].end()) return PROXY_NOT_FOUND;
      dynamic_cast<Image_List&>(value) = it->second;
      return PROXY_FOUND;
    }
    if (const ID* id = dynamic_cast<const ID*>(&key)) {
      uint64_t i = (uint64_t)id->id() - first_id;
      if (!codes || i >= n) return PROXY_NOT_FOUND;
      dynamic_cast<BinaryCode&>(value).set_code((const char*)codes + i * nbytes, nbytes);
      return PROXY_FOUND;
    }
    return PROXY_NOT_FOUND;
  }
  int put(const google::protobuf::Message& key, const google::protobuf::Message& value) {
    const HashIndex* h = dynamic_cast<const HashIndex*>(&key);
    const Image_List* l = dynamic_cast<const Image_List*>(&value);
    if (!h || !l) return PROXY_PUT_FAIL;
    if (h->table_id() >= tables.size()) tables.resize(h->table_id() + 1);
    tables[h->table_id()][h->index()] = *l;
    return PROXY_PUT_DONE;
  }
  int contain(const google::protobuf::Message&) { return 0; }
  int init(const char*) { return 0; }
  void close() {}
};
MemProxy g_mem;
}  // namespace

extern "C" {

// Empties the store (all tables and the main table).
int ref_reset(void) { vc_shim_kv_clear(); return 0; }

// build-tables: runs the reference's main() of src/build_hash_tables.cc once
// per rank (= per table) on a raw code file (record = bits/8 bytes, id = ordinal).
// Its progress printf (build_hash_tables.cc:50-51) is swallowed.
int ref_build_tables(const char* code_file, int binary_bits, int n_tables) {
  int rc = 0;
  capture_stdout([&] { rc = run_main_per_rank(ref_build_tables_main, code_file, binary_bits, n_tables); });
  return rc;
}

// integrity-check: the reference's main() of src/integrity_check.cc, per rank.
// A failed check is an assert() in the reference (integrity_check.cc:61), i.e.
// the process aborts; run it in a child process if that matters.
int ref_integrity_check(const char* code_file, int binary_bits, int n_tables) {
  int rc = 0;
  capture_stdout([&] { rc = run_main_per_rank(ref_integrity_check_main, code_file, binary_bits, n_tables); });
  return rc;
}

// Fills the main table id -> code through the reference's proxy (deviation 2).
int ref_put_main_table(const uint8_t* codes, uint64_t n, int nbytes, uint32_t first_id) {
  Proxy* p = new_proxy();
  ID key; BinaryCode val;
  for (uint64_t i = 0; i < n; ++i) {
    key.set_id(first_id + (uint32_t)i);
    val.set_code((const char*)codes + i * nbytes, nbytes);
    if (p->put(key, val) != PROXY_PUT_DONE) return 1;
  }
  p->close(); delete p;
  return 0;
}

// BaseProxy::get(HashIndex) through the reference's PilafProxy: returns
// PROXY_FOUND(0) / PROXY_NOT_FOUND(1); members in stored order.
int ref_bucket_get(uint32_t table, uint32_t index, int nbytes, uint32_t* ids, uint8_t* codes, uint32_t cap, uint32_t* n_out) {
  Proxy* p = new_proxy();
  HashIndex key; Image_List val;
  key.set_table_id(table); key.set_index(index);
  int rc = p->get(key, val);
  uint32_t n = 0;
  if (rc == PROXY_FOUND) {
    n = (uint32_t)val.images_size();
    for (uint32_t i = 0; i < n && i < cap; ++i) {
      ids[i] = val.images(i).id();
      memcpy(codes + (size_t)i * nbytes, val.images(i).code().data(), nbytes);
    }
  }
  *n_out = n;
  p->close(); delete p;
  return rc;
}

// SearchWorker::find for a batch of queries; n_tables ranks (threads), each with
// its own SearchWorker + PilafProxy, as `mpirun -n n_tables` would have.  Rank 0
// holds the result list (search_worker.cc:210-216), copied out in the
// reference's order (descending distance).  out_* are [nq][k]; unused slots are
// left untouched; out_counts[q] = list length.
int ref_mih_search(const uint8_t* queries, int nq, int nbytes, int n_tables, int k, int approximate, int image_count,
                   uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_counts,
                   uint32_t* out_radius, uint64_t* out_sub_reads) {
  if (nbytes % n_tables != 0) return 1;     // search_worker.cc:75 would assert
  vc_shim_mpi_set_world(n_tables);
  std::vector<std::thread> th;
  for (int r = 0; r < n_tables; ++r) {
    th.emplace_back([=] {
      vc_shim_mpi_set_rank(r);
      mpi_coordinator coord;
      Proxy* proxy = new_proxy();
      SearchWorker worker(&coord, proxy, image_count);
      for (int q = 0; q < nq; ++q) {
        std::list<SearchWorker::search_result_st> res =
            worker.find((const char*)queries + (size_t)q * nbytes, nbytes, k, approximate != 0);
        uint64_t n_main, n_sub, n_local; uint32_t radius;
        worker.get_stat(n_main, n_sub, n_local, radius);
        if (out_sub_reads) out_sub_reads[(size_t)q * n_tables + r] = n_sub;
        if (r == 0) {
          uint32_t i = 0;
          for (auto it = res.begin(); it != res.end(); ++it, ++i) {
            out_ids[(size_t)q * k + i] = it->image_id;
            out_dists[(size_t)q * k + i] = it->dist;
          }
          out_counts[q] = i;
          if (out_radius) out_radius[q] = radius;
        }
      }
      proxy->close();
      delete proxy;
    });
  }
  for (auto& t : th) t.join();
  vc_shim_mpi_set_world(1);
  vc_shim_mpi_set_rank(0);
  return 0;
}

// linear_search.cc:39-64 for one query.  The function prints its result
// ("Find image with id=%d and hamming_dist=%d", descending distance); the
// print-out is captured and parsed.  Needs the main table (ref_put_main_table).
int ref_linear_search(const uint8_t* query, int nbytes, int k, uint32_t image_count,
                      uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_count) {
  Proxy* p = new_proxy();
  proxy_clt = p;
  search_code.assign((const char*)query, nbytes);
  image_total = (int)image_count;
  std::string text = capture_stdout([&] { search_K_nearest_neighbors(k); });
  proxy_clt = 0;
  p->close(); delete p;
  uint32_t n = 0;
  const char* s = text.c_str();
  while (*s) {
    int id, dist;
    if (sscanf(s, "Find image with id=%d and hamming_dist=%d", &id, &dist) == 2) {
      if ((int)n < k) { out_ids[n] = (uint32_t)id; out_dists[n] = (uint32_t)dist; }
      ++n;
    }
    const char* nl = strchr(s, '\n');
    if (!nl) break;
    s = nl + 1;
  }
  *out_count = n;
  return 0;
}

uint64_t ref_kv_entries(void) { return vc_shim_kv_size(); }

// ---- MemProxy-backed variants (CPU baseline legs; same reference code on top) ----

// Direct table fill: key = the reference's binaryToInt (Pilaf/image_tools.h:12-18)
// on substring t, members appended in id order - the end state that
// build_hash_tables.cc:38-64 reaches through get/append/put.  tests/ check it
// against the tables ref_build_tables() produces.
int ref_mem_build(const uint8_t* codes, uint64_t n, int nbytes, int n_tables, uint32_t first_id) {
  g_mem.codes = codes; g_mem.n = n; g_mem.nbytes = nbytes; g_mem.first_id = first_id;
  g_mem.tables.clear();
  g_mem.tables.resize(n_tables);
  if (n_tables == 0) return 0;
  int sub = nbytes / n_tables;
  for (int t = 0; t < n_tables; ++t) {
    auto& tab = g_mem.tables[t];
    for (uint64_t i = 0; i < n; ++i) {
      const char* c = (const char*)codes + i * nbytes;
      uint32_t index = binaryToInt(c + t * sub, sub);
      ID_Code_Pair* pr = tab[index].add_images();
      pr->set_id(first_id + (uint32_t)i);
      pr->set_code(c, nbytes);
    }
  }
  return 0;
}

int ref_mem_bucket_get(uint32_t table, uint32_t index, int nbytes, uint32_t* ids, uint8_t* codes, uint32_t cap, uint32_t* n_out) {
  HashIndex key; Image_List val;
  key.set_table_id(table); key.set_index(index);
  int rc = g_mem.get(key, val);
  uint32_t n = 0;
  if (rc == PROXY_FOUND) {
    n = (uint32_t)val.images_size();
    for (uint32_t i = 0; i < n && i < cap; ++i) {
      ids[i] = val.images(i).id();
      memcpy(codes + (size_t)i * nbytes, val.images(i).code().data(), nbytes);
    }
  }
  *n_out = n;
  return rc;
}

// SearchWorker::find over the MemProxy tables; otherwise identical to ref_mih_search.
int ref_mem_mih_search(const uint8_t* queries, int nq, int nbytes, int n_tables, int k, int approximate, int image_count,
                       uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_counts,
                       uint32_t* out_radius, uint64_t* out_sub_reads) {
  if (nbytes % n_tables != 0) return 1;
  vc_shim_mpi_set_world(n_tables);
  std::vector<std::thread> th;
  for (int r = 0; r < n_tables; ++r) {
    th.emplace_back([=] {
      vc_shim_mpi_set_rank(r);
      mpi_coordinator coord;
      SearchWorker worker(&coord, &g_mem, image_count);
      for (int q = 0; q < nq; ++q) {
        std::list<SearchWorker::search_result_st> res =
            worker.find((const char*)queries + (size_t)q * nbytes, nbytes, k, approximate != 0);
        uint64_t n_main, n_sub, n_local; uint32_t radius;
        worker.get_stat(n_main, n_sub, n_local, radius);
        if (out_sub_reads) out_sub_reads[(size_t)q * n_tables + r] = n_sub;
        if (r == 0) {
          uint32_t i = 0;
          for (auto it = res.begin(); it != res.end(); ++it, ++i) {
            out_ids[(size_t)q * k + i] = it->image_id;
            out_dists[(size_t)q * k + i] = it->dist;
          }
          out_counts[q] = i;
          if (out_radius) out_radius[q] = radius;
        }
      }
    });
  }
  for (auto& t : th) t.join();
  vc_shim_mpi_set_world(1);
  vc_shim_mpi_set_rank(0);
  return 0;
}

// The reference's ImageBitmap (src/bitmap.cc:22-38, unmodified) over a caller-owned word array: op 0 = get_idx, 1 = set_idx,
// 2 = reset_idx for every bit index of the list, in order; get results go to out (0 / 1).  The class frees what it is
// given (bitmap.cc:16-21), so it works on a malloc'ed copy that is copied back.
int ref_bitmap_ops(uint32_t* words, uint64_t n_bytes, const uint64_t* bits, const int* ops, uint64_t n_ops, uint8_t* out) {
  void* copy = malloc(n_bytes);
  if (!copy) return 1;
  memcpy(copy, words, n_bytes);
  {
    ImageBitmap bmp(n_bytes, copy);
    for (uint64_t i = 0; i < n_ops; ++i) {
      if (ops[i] == 0) out[i] = bmp.get_idx(bits[i]) ? 1 : 0;
      else if (ops[i] == 1) bmp.set_idx(bits[i]);
      else bmp.reset_idx(bits[i]);
    }
    memcpy(words, bmp.data(), n_bytes);
  }                                                      // ~ImageBitmap frees the copy
  return 0;
}

// Fixed-radius search (BASELINE config C5: radii 0 .. max_radius, no stop rule, the k best of what was found).  The reference
// has no such entry point, but its pieces are reachable from a subclass (the members are protected, search_worker.h:35-63):
// candidate generation is the reference's own search_R_neighbors / enumerate_entry (probe order, bucket reads, distances)
// and its gather_vectors (rank-order concatenation); only the master's bookkeeping between them - first-seen-wins
// de-duplication and the size-k max-heap with strict-less replacement, search_worker.cc:183-197 - is written out again
// here, without the stop test of :204.
extern "C++" bool operator<(const SearchWorker::search_result_st& a, const SearchWorker::search_result_st& b);   // src/search_worker.cc:15, the reference's own
struct FixedRadiusWorker : SearchWorker {
  FixedRadiusWorker(mpi_coordinator* c, Proxy* p, int total) : SearchWorker(c, p, total) {}
  std::list<search_result_st> find_fixed(const char* binary_code, size_t nbytes, int knn, int max_radius) {
    knn_ = knn; knn_found_.clear(); result_.clear();                         // find(), search_worker.cc:67-76
    n_main_reads_ = 0; n_sub_reads_ = 0; n_local_reads_ = 0;
    n_local_bytes_ = (int)(nbytes / coord_->get_size());
    std::string query_code(binary_code, nbytes);
    std::string local = query_code.substr((size_t)coord_->get_rank() * n_local_bytes_, n_local_bytes_);
    const uint32_t search_index = binaryToInt(local.c_str(), n_local_bytes_);
    std::priority_queue<search_result_st> qmax;
    for (int radius = 0; radius <= max_radius && radius <= n_local_bytes_ * 8; ++radius) {
      std::vector<uint64_t> cand;
      search_R_neighbors(query_code, radius, search_index, cand);
      std::vector<uint64_t> all = coord_->gather_vectors(cand);
      if (!coord_->is_master()) continue;
      for (size_t i = 0; i < all.size(); ++i) {
        const uint32_t id = (uint32_t)(all[i] & 0xffffffffu);
        if (knn_found_.find((int)id) != knn_found_.end()) continue;
        knn_found_[(int)id] = 1;
        search_result_st item; item.image_id = id; item.dist = (uint32_t)(all[i] >> 32);
        if ((int)qmax.size() < knn_) qmax.push(item);
        else if (qmax.top().dist > item.dist) { qmax.pop(); qmax.push(item); }
      }
    }
    while (!qmax.empty()) { result_.push_back(qmax.top()); qmax.pop(); }
    return result_;
  }
  uint64_t sub_reads() const { return n_sub_reads_; }
};

int ref_mem_fixed_radius_search(const uint8_t* queries, int nq, int nbytes, int n_tables, int k, int max_radius, int image_count,
                                uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_counts, uint64_t* out_sub_reads) {
  if (nbytes % n_tables != 0) return 1;
  vc_shim_mpi_set_world(n_tables);
  std::vector<std::thread> th;
  for (int r = 0; r < n_tables; ++r) {
    th.emplace_back([=] {
      vc_shim_mpi_set_rank(r);
      mpi_coordinator coord;
      FixedRadiusWorker worker(&coord, &g_mem, image_count);
      for (int q = 0; q < nq; ++q) {
        std::list<SearchWorker::search_result_st> res = worker.find_fixed((const char*)queries + (size_t)q * nbytes, nbytes, k, max_radius);
        if (out_sub_reads) out_sub_reads[(size_t)q * n_tables + r] = worker.sub_reads();
        if (r == 0) {
          uint32_t i = 0;
          for (auto it = res.begin(); it != res.end(); ++it, ++i) { out_ids[(size_t)q * k + i] = it->image_id; out_dists[(size_t)q * k + i] = it->dist; }
          out_counts[q] = i;
        }
      }
    });
  }
  for (auto& t : th) t.join();
  vc_shim_mpi_set_world(1);
  vc_shim_mpi_set_rank(0);
  return 0;
}

// linear_search.cc:39-64 over a caller-owned code array, nq queries fanned out
// over n_procs forked children (the reference function is single-threaded and
// keeps its state in globals, so processes - not threads - are the only way to
// use several cores without touching it).  Outputs are [nq][k].
int ref_mem_linear_search(const uint8_t* codes, uint64_t n, int nbytes, const uint8_t* queries, int nq, int k,
                          int n_procs, uint32_t* out_ids, uint32_t* out_dists, uint32_t* out_counts) {
  g_mem.codes = codes; g_mem.n = n; g_mem.nbytes = nbytes; g_mem.first_id = 0;
  if (n_procs < 1) n_procs = 1;
  if (n_procs > nq) n_procs = nq;
  size_t words = (size_t)nq * k;
  size_t bytes = (2 * words + nq) * sizeof(uint32_t);
  uint32_t* shm = out_ids;
  bool forked = n_procs > 1;
  if (forked) {
    shm = (uint32_t*)mmap(0, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (shm == MAP_FAILED) return 1;
  }
  uint32_t* ids = forked ? shm : out_ids;
  uint32_t* dists = forked ? shm + words : out_dists;
  uint32_t* counts = forked ? shm + 2 * words : out_counts;
  auto run_range = [&](int q0, int q1) {
    for (int q = q0; q < q1; ++q) {
      proxy_clt = &g_mem;
      search_code.assign((const char*)queries + (size_t)q * nbytes, nbytes);
      image_total = (int)n;
      std::string text = capture_stdout([&] { search_K_nearest_neighbors(k); });
      uint32_t c = 0;
      const char* s = text.c_str();
      while (*s) {
        int id, dist;
        if (sscanf(s, "Find image with id=%d and hamming_dist=%d", &id, &dist) == 2) {
          if ((int)c < k) { ids[(size_t)q * k + c] = (uint32_t)id; dists[(size_t)q * k + c] = (uint32_t)dist; }
          ++c;
        }
        const char* nl = strchr(s, '\n');
        if (!nl) break;
        s = nl + 1;
      }
      counts[q] = c;
    }
    proxy_clt = 0;
  };
  if (!forked) { run_range(0, nq); return 0; }
  std::vector<pid_t> kids;
  for (int p = 0; p < n_procs; ++p) {
    int q0 = (int)((long long)nq * p / n_procs), q1 = (int)((long long)nq * (p + 1) / n_procs);
    pid_t pid = fork();
    if (pid < 0) return 1;
    if (pid == 0) { run_range(q0, q1); _exit(0); }
    kids.push_back(pid);
  }
  int rc = 0;
  for (pid_t pid : kids) { int st = 0; waitpid(pid, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 1; }
  memcpy(out_ids, shm, words * sizeof(uint32_t));
  memcpy(out_dists, shm + words, words * sizeof(uint32_t));
  memcpy(out_counts, shm + 2 * words, nq * sizeof(uint32_t));
  munmap(shm, bytes);
  return rc;
}

}  // extern "C"
