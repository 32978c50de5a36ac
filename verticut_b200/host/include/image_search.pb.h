// Messages of the reference's KV schema (src/image_search.proto:3-27), written by hand because protoc /
// libprotobuf are not part of this build: ID, BinaryCode, HashIndex, ID_Code_Pair, Image_List, ImageList.
// Same accessor names as protoc would generate for the fields the reference uses, and
// SerializeToString / ParseFromString speak the proto2 wire format of those messages
//   ID            08 varint(id)
//   BinaryCode    0A varint(len) bytes
//   HashIndex     08 varint(table_id) 10 varint(index)
//   ID_Code_Pair  08 varint(id) 12 varint(len) code
//   Image_List    { 0A varint(len(pair)) pair }*
//   ImageList     { 08 varint(image) }*
// so bytes written here are readable by a real protobuf build of the reference and vice versa.
#ifndef VERTICUT_B200_IMAGE_SEARCH_PB_H
#define VERTICUT_B200_IMAGE_SEARCH_PB_H

#include <stdint.h>
#include <string>
#include <vector>

namespace google {
namespace protobuf {

class Message {
 public:
  virtual ~Message() {}
  virtual bool SerializeToString(std::string* out) const = 0;
  virtual bool ParseFromString(const std::string& in) = 0;
  virtual void Clear() = 0;
};

namespace wire {
inline void put_varint(std::string* s, uint64_t v) {
  for (; v >= 0x80; v >>= 7) s->push_back(static_cast<char>((v & 0x7f) | 0x80));
  s->push_back(static_cast<char>(v));
}
inline bool get_varint(const std::string& s, size_t* pos, uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; *pos < s.size() && shift < 64; shift += 7) {
    const uint8_t b = static_cast<uint8_t>(s[(*pos)++]);
    r |= static_cast<uint64_t>(b & 0x7f) << shift;
    if (!(b & 0x80)) { *v = r; return true; }
  }
  return false;
}
inline void put_bytes(std::string* s, uint8_t tag, const std::string& b) {
  s->push_back(static_cast<char>(tag));
  put_varint(s, b.size());
  s->append(b);
}
inline bool get_tag(const std::string& s, size_t* pos, uint8_t tag) {
  if (*pos >= s.size() || static_cast<uint8_t>(s[*pos]) != tag) return false;
  ++*pos;
  return true;
}
inline bool get_bytes(const std::string& s, size_t* pos, std::string* out) {
  uint64_t len;
  if (!get_varint(s, pos, &len) || len > s.size() - *pos) return false;
  out->assign(s, *pos, static_cast<size_t>(len));
  *pos += static_cast<size_t>(len);
  return true;
}
}  // namespace wire
}  // namespace protobuf
}  // namespace google

class ID : public google::protobuf::Message {
 public:
  ID() : id_(0) {}
  void set_id(uint32_t v) { id_ = v; }
  uint32_t id() const { return id_; }
  void Clear() { id_ = 0; }
  bool SerializeToString(std::string* out) const {
    out->clear();
    out->push_back(0x08);
    google::protobuf::wire::put_varint(out, id_);
    return true;
  }
  bool ParseFromString(const std::string& in) {
    size_t p = 0;
    uint64_t v;
    if (!google::protobuf::wire::get_tag(in, &p, 0x08) || !google::protobuf::wire::get_varint(in, &p, &v)) return false;
    id_ = static_cast<uint32_t>(v);
    return p == in.size();
  }

 private:
  uint32_t id_;
};

class BinaryCode : public google::protobuf::Message {
 public:
  void set_code(const char* p, size_t n) { code_.assign(p, n); }
  void set_code(const std::string& s) { code_ = s; }
  const std::string& code() const { return code_; }
  void Clear() { code_.clear(); }
  bool SerializeToString(std::string* out) const {
    out->clear();
    google::protobuf::wire::put_bytes(out, 0x0A, code_);
    return true;
  }
  bool ParseFromString(const std::string& in) {
    size_t p = 0;
    return google::protobuf::wire::get_tag(in, &p, 0x0A) && google::protobuf::wire::get_bytes(in, &p, &code_) && p == in.size();
  }

 private:
  std::string code_;
};

class HashIndex : public google::protobuf::Message {
 public:
  HashIndex() : table_id_(0), index_(0) {}
  void set_table_id(uint32_t v) { table_id_ = v; }
  void set_index(uint32_t v) { index_ = v; }
  uint32_t table_id() const { return table_id_; }
  uint32_t index() const { return index_; }
  void Clear() { table_id_ = index_ = 0; }
  bool SerializeToString(std::string* out) const {
    out->clear();
    out->push_back(0x08);
    google::protobuf::wire::put_varint(out, table_id_);
    out->push_back(0x10);
    google::protobuf::wire::put_varint(out, index_);
    return true;
  }
  bool ParseFromString(const std::string& in) {
    size_t p = 0;
    uint64_t a, b;
    if (!google::protobuf::wire::get_tag(in, &p, 0x08) || !google::protobuf::wire::get_varint(in, &p, &a)) return false;
    if (!google::protobuf::wire::get_tag(in, &p, 0x10) || !google::protobuf::wire::get_varint(in, &p, &b)) return false;
    table_id_ = static_cast<uint32_t>(a);
    index_ = static_cast<uint32_t>(b);
    return p == in.size();
  }

 private:
  uint32_t table_id_, index_;
};

class ID_Code_Pair : public google::protobuf::Message {
 public:
  ID_Code_Pair() : id_(0) {}
  void set_id(uint32_t v) { id_ = v; }
  uint32_t id() const { return id_; }
  void set_code(const char* p, size_t n) { code_.assign(p, n); }
  void set_code(const std::string& s) { code_ = s; }
  const std::string& code() const { return code_; }
  void Clear() { id_ = 0; code_.clear(); }
  void AppendToString(std::string* out) const {
    out->push_back(0x08);
    google::protobuf::wire::put_varint(out, id_);
    google::protobuf::wire::put_bytes(out, 0x12, code_);
  }
  bool SerializeToString(std::string* out) const {
    out->clear();
    AppendToString(out);
    return true;
  }
  bool ParseFromString(const std::string& in) {
    size_t p = 0;
    uint64_t v;
    if (!google::protobuf::wire::get_tag(in, &p, 0x08) || !google::protobuf::wire::get_varint(in, &p, &v)) return false;
    id_ = static_cast<uint32_t>(v);
    return google::protobuf::wire::get_tag(in, &p, 0x12) && google::protobuf::wire::get_bytes(in, &p, &code_) && p == in.size();
  }

 private:
  uint32_t id_;
  std::string code_;
};

class Image_List : public google::protobuf::Message {
 public:
  int images_size() const { return static_cast<int>(images_.size()); }
  const ID_Code_Pair& images(int i) const { return images_[i]; }
  ID_Code_Pair* add_images() {
    images_.push_back(ID_Code_Pair());
    return &images_.back();
  }
  void clear_images() { images_.clear(); }
  void Clear() { images_.clear(); }
  bool SerializeToString(std::string* out) const {
    out->clear();
    std::string one;
    for (size_t i = 0; i < images_.size(); ++i) {
      one.clear();
      images_[i].AppendToString(&one);
      google::protobuf::wire::put_bytes(out, 0x0A, one);
    }
    return true;
  }
  bool ParseFromString(const std::string& in) {
    images_.clear();
    size_t p = 0;
    std::string one;
    while (p < in.size()) {
      if (!google::protobuf::wire::get_tag(in, &p, 0x0A) || !google::protobuf::wire::get_bytes(in, &p, &one)) return false;
      ID_Code_Pair pr;
      if (!pr.ParseFromString(one)) return false;
      images_.push_back(pr);
    }
    return true;
  }

 private:
  std::vector<ID_Code_Pair> images_;
};

class ImageList : public google::protobuf::Message {
 public:
  int images_size() const { return static_cast<int>(images_.size()); }
  uint32_t images(int i) const { return images_[i]; }
  void add_images(uint32_t v) { images_.push_back(v); }
  void clear_images() { images_.clear(); }
  void Clear() { images_.clear(); }
  bool SerializeToString(std::string* out) const {
    out->clear();
    for (size_t i = 0; i < images_.size(); ++i) {
      out->push_back(0x08);
      google::protobuf::wire::put_varint(out, images_[i]);
    }
    return true;
  }
  bool ParseFromString(const std::string& in) {
    images_.clear();
    size_t p = 0;
    uint64_t v;
    while (p < in.size()) {
      if (!google::protobuf::wire::get_tag(in, &p, 0x08) || !google::protobuf::wire::get_varint(in, &p, &v)) return false;
      images_.push_back(static_cast<uint32_t>(v));
    }
    return true;
  }

 private:
  std::vector<uint32_t> images_;
};

#endif
