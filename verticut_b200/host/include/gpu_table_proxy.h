// GpuTableProxy - the BaseProxy backend that keeps the MIH tables in the HBM of one B200 instead of a
// memcached / redis / Pilaf cluster (the reference's src/memcached_proxy.h, src/redis_proxy.h,
// src/pilaf_proxy.h).  Same interface (base_proxy.h), same key/value kinds as the reference uses:
//   (HashIndex{table_id, index} -> Image_List)   src/search_worker.cc:246, src/build_hash_tables.cc:48,57,63
//   (ID{id}                     -> BinaryCode)   src/linear_search.cc:45-46
// Everything below is host glue over the C ABI (include/verticut_gpu.h); no arithmetic of the path lives here.
//
// Life cycle.  init(filename) reads the "server list" - here one CUDA device ordinal per line (first line
// used; NULL or an empty file = device 0) - and creates the index.  While codes are being put, the proxy is
// in LOADING state: put() stores values exactly like a KV store would (so the reference's read-modify-write
// build loop, get -> append -> put, works unchanged), and finalize() - called explicitly, or implicitly by the
// first search - moves every (id, code) pair to the GPU and builds the CSR tables there.  load_codes() is the
// fast path that skips the KV emulation: raw records straight to HBM.
// Like the reference's proxies the object is single-threaded (src/pilaf_proxy.h:19-21).
#ifndef VERTICUT_B200_GPU_TABLE_PROXY_H
#define VERTICUT_B200_GPU_TABLE_PROXY_H

#include <stdint.h>
#include <map>
#include <string>
#include <unordered_map>

#include "base_proxy.h"
#include "image_search.pb.h"
#include "verticut_gpu.h"

class GpuTableProxy : public BaseProxy<google::protobuf::Message, google::protobuf::Message> {
 public:
  GpuTableProxy(int binary_bits, int n_tables, uint32_t first_id = 0);
  ~GpuTableProxy();

  int get(const google::protobuf::Message& key, google::protobuf::Message& value);
  int put(const google::protobuf::Message& key, const google::protobuf::Message& value);
  int contain(const google::protobuf::Message& key);   // 0, as in every backend of the reference ("Not implemented yet")
  int init(const char* filename);
  void close();

  // fast path: n raw records (binary_bits/8 bytes each), ids continue from the last one
  int load_codes(const void* codes, uint64_t n);
  int load_code_file(const char* path, uint64_t max_codes = 0);
  // uploads what was put() and (re)builds the tables; 0 on success
  int finalize();
  // persistence: the reference's tables outlive the builder process in the KV servers; here they can be written
  // to a file by build-tables and read back by the search tools (vc_index_save / vc_index_load)
  int save(const char* path);
  int load(const char* path);          // replaces init() + load_codes() + finalize()

  vc_index* handle() { return ix_; }
  uint64_t size() const;
  int code_bytes() const { return bits_ / 8; }
  int n_tables() const { return tables_; }
  const char* last_error() const { return vc_last_error(); }
  // what the config file of init() said about the origin of the tables ("index <file>" / "codes <file>" lines); empty if nothing
  const std::string& config_index_path() const { return index_path_; }
  const std::string& config_codes_path() const { return codes_path_; }

 private:
  GpuTableProxy(const GpuTableProxy&);
  int bits_, tables_, device_;
  uint32_t first_id_;
  vc_index* ix_;
  std::string index_path_, codes_path_;
  bool dirty_;                                             // codes added since the last build
  std::map<uint32_t, std::string> staged_codes_;           // id -> code, from put(); drained by finalize()
  std::unordered_map<uint64_t, Image_List> staged_lists_;  // (table << 32 | index) -> value, LOADING state only
};

#endif
