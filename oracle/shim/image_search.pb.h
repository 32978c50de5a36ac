// TEST INFRASTRUCTURE ONLY (oracle).  Hand-written stand-in for the header
// protoc would generate from the reference's src/image_search.proto:3-27, so
// the reference sources compile here without protoc/libprotobuf.  It provides
// exactly the accessors the reference calls (set_id/id, set_code/code,
// set_table_id/set_index, images_size/images/add_images/clear_images) and a
// SerializeToString/ParseFromString pair that speaks the proto2 wire format
// of those messages, so the reference's own proxies
// (pilaf_proxy.h:41-62) round-trip through bytes just as they would with
// real protobuf.  No arithmetic of the hot path lives here.
#ifndef VC_ORACLE_SHIM_IMAGE_SEARCH_PB_H
#define VC_ORACLE_SHIM_IMAGE_SEARCH_PB_H

#include <stdint.h>
#include <string.h>
#include <assert.h>
#include <stdlib.h>
#include <iostream>
#include <list>
#include <map>
#include <queue>
#include <string>
#include <vector>

namespace google {
namespace protobuf {

class Message {
 public:
  virtual ~Message() {}
  virtual bool SerializeToString(std::string* out) const = 0;
  virtual bool ParseFromString(const std::string& in) = 0;
};

namespace shim_wire {
inline void put_varint(std::string* s, uint64_t v) {
  while (v >= 0x80) { s->push_back((char)((v & 0x7f) | 0x80)); v >>= 7; }
  s->push_back((char)v);
}
inline bool get_varint(const std::string& s, size_t* pos, uint64_t* v) {
  uint64_t r = 0; int shift = 0;
  while (*pos < s.size() && shift < 64) {
    uint8_t b = (uint8_t)s[(*pos)++];
    r |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) { *v = r; return true; }
    shift += 7;
  }
  return false;
}
inline bool get_bytes(const std::string& s, size_t* pos, std::string* out) {
  uint64_t len;
  if (!get_varint(s, pos, &len) || *pos + len > s.size()) return false;
  out->assign(s, *pos, (size_t)len);
  *pos += (size_t)len;
  return true;
}
}  // namespace shim_wire
}  // namespace protobuf
}  // namespace google

// message ID { required uint32 id = 1; }
class ID : public google::protobuf::Message {
  uint32_t id_;
 public:
  ID() : id_(0) {}
  void set_id(uint32_t v) { id_ = v; }
  uint32_t id() const { return id_; }
  bool SerializeToString(std::string* out) const {
    out->clear(); out->push_back(0x08);
    google::protobuf::shim_wire::put_varint(out, id_); return true;
  }
  bool ParseFromString(const std::string& in) {
    size_t p = 0; uint64_t v;
    if (in.size() < 2 || in[p++] != 0x08 || !google::protobuf::shim_wire::get_varint(in, &p, &v)) return false;
    id_ = (uint32_t)v; return true;
  }
};

// message BinaryCode { required bytes code = 1; }
class BinaryCode : public google::protobuf::Message {
  std::string code_;
 public:
  void set_code(const char* p, size_t n) { code_.assign(p, n); }
  void set_code(const std::string& s) { code_ = s; }
  const std::string& code() const { return code_; }
  bool SerializeToString(std::string* out) const {
    out->clear(); out->push_back(0x0A);
    google::protobuf::shim_wire::put_varint(out, code_.size()); out->append(code_); return true;
  }
  bool ParseFromString(const std::string& in) {
    size_t p = 0;
    if (in.empty() || in[p++] != 0x0A) return false;
    return google::protobuf::shim_wire::get_bytes(in, &p, &code_);
  }
};

// message HashIndex { required uint32 table_id = 1; required uint32 index = 2; }
class HashIndex : public google::protobuf::Message {
  uint32_t table_id_, index_;
 public:
  HashIndex() : table_id_(0), index_(0) {}
  void set_table_id(uint32_t v) { table_id_ = v; }
  void set_index(uint32_t v) { index_ = v; }
  uint32_t table_id() const { return table_id_; }
  uint32_t index() const { return index_; }
  bool SerializeToString(std::string* out) const {
    out->clear();
    out->push_back(0x08); google::protobuf::shim_wire::put_varint(out, table_id_);
    out->push_back(0x10); google::protobuf::shim_wire::put_varint(out, index_);
    return true;
  }
  bool ParseFromString(const std::string& in) {
    size_t p = 0; uint64_t v;
    if (p >= in.size() || in[p++] != 0x08 || !google::protobuf::shim_wire::get_varint(in, &p, &v)) return false;
    table_id_ = (uint32_t)v;
    if (p >= in.size() || in[p++] != 0x10 || !google::protobuf::shim_wire::get_varint(in, &p, &v)) return false;
    index_ = (uint32_t)v; return true;
  }
};

// message ID_Code_Pair { required uint32 id = 1; required bytes code = 2; }
class ID_Code_Pair : public google::protobuf::Message {
  uint32_t id_;
  std::string code_;
 public:
  ID_Code_Pair() : id_(0) {}
  void set_id(uint32_t v) { id_ = v; }
  uint32_t id() const { return id_; }
  void set_code(const char* p, size_t n) { code_.assign(p, n); }
  void set_code(const std::string& s) { code_ = s; }
  const std::string& code() const { return code_; }
  void AppendTo(std::string* out) const {
    out->push_back(0x08); google::protobuf::shim_wire::put_varint(out, id_);
    out->push_back(0x12); google::protobuf::shim_wire::put_varint(out, code_.size()); out->append(code_);
  }
  bool SerializeToString(std::string* out) const { out->clear(); AppendTo(out); return true; }
  bool ParseFromString(const std::string& in) {
    size_t p = 0; uint64_t v;
    if (p >= in.size() || in[p++] != 0x08 || !google::protobuf::shim_wire::get_varint(in, &p, &v)) return false;
    id_ = (uint32_t)v;
    if (p >= in.size() || in[p++] != 0x12) return false;
    return google::protobuf::shim_wire::get_bytes(in, &p, &code_);
  }
};

// message Image_List { repeated ID_Code_Pair images = 1; }
class Image_List : public google::protobuf::Message {
  std::vector<ID_Code_Pair> images_;
 public:
  int images_size() const { return (int)images_.size(); }
  const ID_Code_Pair& images(int i) const { return images_[i]; }
  ID_Code_Pair* add_images() { images_.push_back(ID_Code_Pair()); return &images_.back(); }
  void clear_images() { images_.clear(); }
  bool SerializeToString(std::string* out) const {
    out->clear();
    std::string tmp;
    for (size_t i = 0; i < images_.size(); ++i) {
      tmp.clear(); images_[i].AppendTo(&tmp);
      out->push_back(0x0A); google::protobuf::shim_wire::put_varint(out, tmp.size()); out->append(tmp);
    }
    return true;
  }
  bool ParseFromString(const std::string& in) {
    images_.clear();
    size_t p = 0; std::string sub;
    while (p < in.size()) {
      if (in[p++] != 0x0A || !google::protobuf::shim_wire::get_bytes(in, &p, &sub)) return false;
      ID_Code_Pair pr;
      if (!pr.ParseFromString(sub)) return false;
      images_.push_back(pr);
    }
    return true;
  }
};

// message ImageList { repeated uint32 images = 1; }  (unused by the reference; kept for completeness)
class ImageList : public google::protobuf::Message {
  std::vector<uint32_t> images_;
 public:
  int images_size() const { return (int)images_.size(); }
  uint32_t images(int i) const { return images_[i]; }
  void add_images(uint32_t v) { images_.push_back(v); }
  void clear_images() { images_.clear(); }
  bool SerializeToString(std::string* out) const {
    out->clear();
    for (size_t i = 0; i < images_.size(); ++i) { out->push_back(0x08); google::protobuf::shim_wire::put_varint(out, images_[i]); }
    return true;
  }
  bool ParseFromString(const std::string& in) {
    images_.clear(); size_t p = 0; uint64_t v;
    while (p < in.size()) {
      if (in[p++] != 0x08 || !google::protobuf::shim_wire::get_varint(in, &p, &v)) return false;
      images_.push_back((uint32_t)v);
    }
    return true;
  }
};

#endif
