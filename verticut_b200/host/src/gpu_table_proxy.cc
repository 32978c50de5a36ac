#include "gpu_table_proxy.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <sstream>
#include <vector>

GpuTableProxy::GpuTableProxy(int binary_bits, int n_tables, uint32_t first_id)
    : bits_(binary_bits), tables_(n_tables), device_(0), first_id_(first_id), ix_(0), dirty_(false) {}

GpuTableProxy::~GpuTableProxy() { close(); }

int GpuTableProxy::init(const char* filename) {
  device_ = 0;
  index_path_.clear(); codes_path_.clear();
  if (filename) {
    std::ifstream fin(filename);
    if (!fin.is_open()) return -1;             // same failure mode as the reference's proxies
    // "server list" of the GPU backend: the first bare number is the CUDA device ordinal; optional lines
    // "index <file>" / "codes <file>" say where this server's tables come from (the reference's servers hold theirs
    // already; tools started with the reference's positional arguments have no other way to name them)
    std::string line;
    bool have_dev = false;
    while (std::getline(fin, line)) {
      std::istringstream ls(line);
      std::string tok;
      if (!(ls >> tok) || tok[0] == '#') continue;
      if (tok == "index" || tok == "codes") {
        std::string path;
        if (ls >> path) (tok == "index" ? index_path_ : codes_path_) = path;
      } else if (!have_dev && (isdigit((unsigned char)tok[0]))) {
        device_ = atoi(tok.c_str());
        have_dev = true;
      }
    }
  }
  if (ix_) { vc_index_destroy(ix_); ix_ = 0; }
  if (vc_index_create(device_, (uint32_t)bits_, (uint32_t)tables_, first_id_, &ix_) != VC_OK) {
    fprintf(stderr, "GpuTableProxy: %s\n", vc_last_error());
    return -1;
  }
  dirty_ = false;
  return 0;
}

void GpuTableProxy::close() {
  if (ix_) { vc_index_destroy(ix_); ix_ = 0; }
  staged_codes_.clear();
  staged_lists_.clear();
}

uint64_t GpuTableProxy::size() const {
  vc_index_info info;
  if (!ix_ || vc_index_get_info(ix_, &info) != VC_OK) return 0;
  return info.n_codes + staged_codes_.size();
}

int GpuTableProxy::load_codes(const void* codes, uint64_t n) {
  if (!ix_) return -1;
  if (vc_index_add(ix_, codes, n) != VC_OK) return -1;
  dirty_ = true;
  return 0;
}

int GpuTableProxy::load_code_file(const char* path, uint64_t max_codes) {
  // raw records, id = ordinal (src/build_hash_tables.cc:40-45,55,69), streamed in 64 MB pieces
  FILE* fh = fopen(path, "rb");
  if (!fh) return -1;
  const size_t rec = (size_t)bits_ / 8, chunk = (64u << 20) / rec;
  std::vector<char> buf(chunk * rec);
  uint64_t total = 0;
  for (;;) {
    size_t want = chunk;
    if (max_codes && total + want > max_codes) want = (size_t)(max_codes - total);
    if (want == 0) break;
    const size_t got = fread(buf.data(), rec, want, fh);
    if (got == 0) break;
    if (load_codes(buf.data(), got) != 0) { fclose(fh); return -1; }
    total += got;
  }
  fclose(fh);
  return 0;
}

int GpuTableProxy::save(const char* path) {
  if (finalize() != 0) return -1;
  return vc_index_save(ix_, path) == VC_OK ? 0 : -1;
}

int GpuTableProxy::load(const char* path) {
  if (ix_) { vc_index_destroy(ix_); ix_ = 0; }
  if (vc_index_load(device_, path, &ix_) != VC_OK) { fprintf(stderr, "GpuTableProxy: %s\n", vc_last_error()); return -1; }
  vc_index_info info;
  vc_index_get_info(ix_, &info);
  bits_ = (int)info.code_bits; tables_ = (int)info.n_tables; first_id_ = info.first_id;
  dirty_ = false;
  return 0;
}

int GpuTableProxy::finalize() {
  if (!ix_) return -1;
  if (!staged_codes_.empty()) {
    // the ids put so far must continue the ids already on the device: ids are file ordinals in the reference
    vc_index_info info;
    if (vc_index_get_info(ix_, &info) != VC_OK) return -1;
    uint64_t expect = (uint64_t)first_id_ + info.n_codes;
    const size_t rec = (size_t)bits_ / 8;
    std::vector<char> flat;
    flat.reserve(staged_codes_.size() * rec);
    for (std::map<uint32_t, std::string>::const_iterator it = staged_codes_.begin(); it != staged_codes_.end(); ++it, ++expect) {
      if (it->first != expect || it->second.size() != rec) {
        fprintf(stderr, "GpuTableProxy: put() ids must be consecutive ordinals with %zu-byte codes (got id %u)\n", rec, it->first);
        return -1;
      }
      flat.insert(flat.end(), it->second.begin(), it->second.end());
    }
    if (vc_index_add(ix_, flat.data(), staged_codes_.size()) != VC_OK) return -1;
    staged_codes_.clear();
    staged_lists_.clear();
    dirty_ = true;
  }
  if (dirty_) {
    if (vc_index_build(ix_) != VC_OK) { fprintf(stderr, "GpuTableProxy: %s\n", vc_last_error()); return -1; }
    dirty_ = false;
  }
  return 0;
}

int GpuTableProxy::put(const google::protobuf::Message& key, const google::protobuf::Message& value) {
  if (!ix_) return PROXY_PUT_FAIL;
  const size_t rec = (size_t)bits_ / 8;
  if (const HashIndex* h = dynamic_cast<const HashIndex*>(&key)) {
    const Image_List* l = dynamic_cast<const Image_List*>(&value);
    if (!l || (int)h->table_id() >= tables_) return PROXY_PUT_FAIL;
    for (int i = 0; i < l->images_size(); ++i) {
      if (l->images(i).code().size() != rec) return PROXY_PUT_FAIL;
      staged_codes_[l->images(i).id()] = l->images(i).code();
    }
    staged_lists_[((uint64_t)h->table_id() << 32) | h->index()] = *l;
    return PROXY_PUT_DONE;
  }
  if (const ID* id = dynamic_cast<const ID*>(&key)) {
    const BinaryCode* c = dynamic_cast<const BinaryCode*>(&value);
    if (!c || c->code().size() != rec) return PROXY_PUT_FAIL;
    staged_codes_[id->id()] = c->code();
    return PROXY_PUT_DONE;
  }
  return PROXY_PUT_FAIL;
}

int GpuTableProxy::get(const google::protobuf::Message& key, google::protobuf::Message& value) {
  if (!ix_) return PROXY_NOT_FOUND;
  const size_t rec = (size_t)bits_ / 8;
  if (const HashIndex* h = dynamic_cast<const HashIndex*>(&key)) {
    Image_List* l = dynamic_cast<Image_List*>(&value);
    if (!l) return PROXY_NOT_FOUND;
    if (!staged_codes_.empty()) {
      // LOADING state: behave like the KV store the build loop expects (value = what was last put)
      std::unordered_map<uint64_t, Image_List>::const_iterator it = staged_lists_.find(((uint64_t)h->table_id() << 32) | h->index());
      if (it == staged_lists_.end()) return PROXY_NOT_FOUND;
      *l = it->second;
      return PROXY_FOUND;
    }
    if (finalize() != 0) return PROXY_NOT_FOUND;
    uint32_t n = 0;
    int rc = vc_bucket_get(ix_, h->table_id(), h->index(), 0, 0, 0, &n);
    if (rc != VC_OK) return PROXY_NOT_FOUND;
    std::vector<uint32_t> ids(n);
    std::vector<char> codes((size_t)n * rec);
    rc = vc_bucket_get(ix_, h->table_id(), h->index(), ids.data(), codes.data(), n, &n);
    if (rc != VC_OK) return PROXY_NOT_FOUND;
    l->clear_images();
    for (uint32_t i = 0; i < n; ++i) {
      ID_Code_Pair* p = l->add_images();
      p->set_id(ids[i]);
      p->set_code(codes.data() + (size_t)i * rec, rec);
    }
    return PROXY_FOUND;
  }
  if (const ID* id = dynamic_cast<const ID*>(&key)) {
    BinaryCode* c = dynamic_cast<BinaryCode*>(&value);
    if (!c) return PROXY_NOT_FOUND;
    std::map<uint32_t, std::string>::const_iterator it = staged_codes_.find(id->id());
    if (it != staged_codes_.end()) { c->set_code(it->second); return PROXY_FOUND; }
    std::vector<char> code(rec);
    if (vc_code_get(ix_, id->id(), code.data()) != VC_OK) return PROXY_NOT_FOUND;
    c->set_code(code.data(), rec);
    return PROXY_FOUND;
  }
  return PROXY_NOT_FOUND;
}

int GpuTableProxy::contain(const google::protobuf::Message&) { return 0; }
