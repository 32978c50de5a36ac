// msgpack_lite - the subset of the MessagePack format the reference's query RPC puts on the wire
// (src/image_search_client.cc:19-34, src/image_search_server.cc:22-50: method name, uint32 / bool / string
// parameters, list<pair<uint32,uint32>> results inside msgpack-rpc's [type, msgid, ...] arrays).
// Hand-written because neither msgpack-c nor msgpack-rpc (the reference's src/Makefile:5-7 dependencies) exist in
// this image.  Header-only; no allocation beyond std::string / std::vector.
//
// Encoder: emits the formats the reference's msgpack 0.5.x peers understand - strings as "raw" (fixraw / raw16 /
// raw32, never str8, which that generation does not know), smallest integer encoding.
// Decoder: accepts every MessagePack type (str8, bin, ext, floats and maps are parsed and kept or skipped), so a
// current msgpack client can talk to the server as well.  parse() works on a byte stream: it reports NEED_MORE
// until a whole object has arrived (msgpack-rpc sends objects back to back with no framing).
#ifndef VERTICUT_B200_MSGPACK_LITE_H
#define VERTICUT_B200_MSGPACK_LITE_H

#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

namespace mp {

// ---- encoder ---------------------------------------------------------------------------------------------------
class Packer {
 public:
  const std::string& bytes() const { return out_; }
  std::string& bytes() { return out_; }

  void pack_nil() { put(0xc0); }
  void pack_bool(bool v) { put(v ? 0xc3 : 0xc2); }
  void pack_uint(uint64_t v) {
    if (v < 128) put((uint8_t)v);
    else if (v <= 0xff) { put(0xcc); be(v, 1); }
    else if (v <= 0xffff) { put(0xcd); be(v, 2); }
    else if (v <= 0xffffffffull) { put(0xce); be(v, 4); }
    else { put(0xcf); be(v, 8); }
  }
  void pack_int(int64_t v) {
    if (v >= 0) return pack_uint((uint64_t)v);
    if (v >= -32) put((uint8_t)(int8_t)v);
    else if (v >= -128) { put(0xd0); be((uint64_t)v, 1); }
    else if (v >= -32768) { put(0xd1); be((uint64_t)v, 2); }
    else if (v >= -2147483648ll) { put(0xd2); be((uint64_t)v, 4); }
    else { put(0xd3); be((uint64_t)v, 8); }
  }
  // "raw" of msgpack 0.5 = "str" of the current spec minus str8
  void pack_str(const std::string& s) { pack_str(s.data(), s.size()); }
  void pack_str(const char* p, size_t n) {
    if (n < 32) put((uint8_t)(0xa0 | n));
    else if (n <= 0xffff) { put(0xda); be(n, 2); }
    else { put(0xdb); be(n, 4); }
    out_.append(p, n);
  }
  void pack_array(size_t n) {
    if (n < 16) put((uint8_t)(0x90 | n));
    else if (n <= 0xffff) { put(0xdc); be(n, 2); }
    else { put(0xdd); be(n, 4); }
  }

 private:
  void put(uint8_t b) { out_.push_back((char)b); }
  void be(uint64_t v, int nbytes) { for (int i = nbytes - 1; i >= 0; --i) put((uint8_t)(v >> (8 * i))); }
  std::string out_;
};

// ---- decoder ---------------------------------------------------------------------------------------------------
struct Value {
  enum Type { NIL, BOOL, UINT, INT, FLOAT, STR, BIN, ARRAY, MAP, EXT };
  Type type = NIL;
  bool b = false;
  uint64_t u = 0;          // UINT
  int64_t i = 0;           // INT (negative values only; non-negative integers are UINT)
  double f = 0;
  std::string s;           // STR / BIN / EXT payload
  std::vector<Value> a;    // ARRAY items; MAP: key0, value0, key1, value1, ...

  bool is_uint() const { return type == UINT; }
  bool is_str() const { return type == STR || type == BIN; }   // old peers send strings as raw, new ones may send bin
  bool is_array() const { return type == ARRAY; }
};

enum ParseStatus { PARSE_OK = 0, PARSE_NEED_MORE = 1, PARSE_ERROR = 2 };

namespace detail {
inline uint64_t rd(const uint8_t* p, int n) { uint64_t v = 0; for (int i = 0; i < n; ++i) v = (v << 8) | p[i]; return v; }

// `budget`: array / map items the whole message may still declare.  A Value is ~100 bytes, so without it a 1 MB message of
// one-byte items would cost ~100 MB per connection to decode; the callers pass what their interface can need.
inline ParseStatus parse_at(const uint8_t* p, size_t n, size_t& pos, Value& v, int depth, size_t& budget) {
  if (depth > 32) return PARSE_ERROR;                       // nothing on this interface nests deeper than 4
  if (pos >= n) return PARSE_NEED_MORE;
  const uint8_t t = p[pos++];
  auto need = [&](size_t k) { return n - pos >= k; };
  auto blob = [&](size_t len, Value::Type ty) -> ParseStatus {
    if (!need(len)) return PARSE_NEED_MORE;
    v.type = ty; v.s.assign((const char*)p + pos, len); pos += len; return PARSE_OK;
  };
  auto items = [&](size_t cnt, Value::Type ty) -> ParseStatus {
    if (cnt > budget) return PARSE_ERROR;                   // more items than the interface ever sends
    budget -= cnt;
    if (cnt > n - pos) return PARSE_NEED_MORE;              // every item takes at least one byte: bounds the allocation
    v.type = ty; v.a.clear(); v.a.resize(cnt);
    for (size_t k = 0; k < cnt; ++k) { ParseStatus st = parse_at(p, n, pos, v.a[k], depth + 1, budget); if (st != PARSE_OK) return st; }
    return PARSE_OK;
  };
  auto lenpfx = [&](int k, uint64_t& len) -> bool { if (!need((size_t)k)) return false; len = rd(p + pos, k); pos += k; return true; };
  uint64_t len = 0;
  if (t < 0x80) { v.type = Value::UINT; v.u = t; return PARSE_OK; }
  if (t >= 0xe0) { v.type = Value::INT; v.i = (int8_t)t; return PARSE_OK; }
  if (t >= 0xa0 && t <= 0xbf) return blob(t & 0x1f, Value::STR);
  if (t >= 0x90 && t <= 0x9f) return items(t & 0x0f, Value::ARRAY);
  if (t >= 0x80 && t <= 0x8f) return items(2 * (size_t)(t & 0x0f), Value::MAP);
  switch (t) {
    case 0xc0: v.type = Value::NIL; return PARSE_OK;
    case 0xc1: return PARSE_ERROR;
    case 0xc2: case 0xc3: v.type = Value::BOOL; v.b = t == 0xc3; return PARSE_OK;
    case 0xc4: case 0xc5: case 0xc6: if (!lenpfx(1 << (t - 0xc4), len)) return PARSE_NEED_MORE; return blob(len, Value::BIN);
    case 0xc7: case 0xc8: case 0xc9: if (!lenpfx(1 << (t - 0xc7), len)) return PARSE_NEED_MORE; if (!need(1)) return PARSE_NEED_MORE; ++pos; return blob(len, Value::EXT);
    case 0xca: { if (!need(4)) return PARSE_NEED_MORE; uint32_t w = (uint32_t)rd(p + pos, 4); pos += 4; float x; memcpy(&x, &w, 4); v.type = Value::FLOAT; v.f = x; return PARSE_OK; }
    case 0xcb: { if (!need(8)) return PARSE_NEED_MORE; uint64_t w = rd(p + pos, 8); pos += 8; double x; memcpy(&x, &w, 8); v.type = Value::FLOAT; v.f = x; return PARSE_OK; }
    case 0xcc: case 0xcd: case 0xce: case 0xcf: {
      const int k = 1 << (t - 0xcc);
      if (!need((size_t)k)) return PARSE_NEED_MORE;
      v.type = Value::UINT; v.u = rd(p + pos, k); pos += k; return PARSE_OK;
    }
    case 0xd0: case 0xd1: case 0xd2: case 0xd3: {
      const int k = 1 << (t - 0xd0);
      if (!need((size_t)k)) return PARSE_NEED_MORE;
      const uint64_t raw = rd(p + pos, k); pos += k;
      const int64_t sv = k == 8 ? (int64_t)raw : (int64_t)(raw << (64 - 8 * k)) >> (64 - 8 * k);   // sign-extend
      if (sv >= 0) { v.type = Value::UINT; v.u = (uint64_t)sv; } else { v.type = Value::INT; v.i = sv; }
      return PARSE_OK;
    }
    case 0xd4: case 0xd5: case 0xd6: case 0xd7: case 0xd8: if (!need(1)) return PARSE_NEED_MORE; ++pos; return blob((size_t)1 << (t - 0xd4), Value::EXT);
    case 0xd9: if (!lenpfx(1, len)) return PARSE_NEED_MORE; return blob(len, Value::STR);
    case 0xda: if (!lenpfx(2, len)) return PARSE_NEED_MORE; return blob(len, Value::STR);
    case 0xdb: if (!lenpfx(4, len)) return PARSE_NEED_MORE; return blob(len, Value::STR);
    case 0xdc: if (!lenpfx(2, len)) return PARSE_NEED_MORE; return items(len, Value::ARRAY);
    case 0xdd: if (!lenpfx(4, len)) return PARSE_NEED_MORE; return items(len, Value::ARRAY);
    case 0xde: if (!lenpfx(2, len)) return PARSE_NEED_MORE; return items(2 * len, Value::MAP);
    case 0xdf: if (!lenpfx(4, len)) return PARSE_NEED_MORE; return items(2 * len, Value::MAP);
  }
  return PARSE_ERROR;
}
}  // namespace detail

// One object from the front of [p, p + n).  PARSE_OK: `consumed` bytes were used.  PARSE_NEED_MORE: the object is
// not complete yet (call again with more bytes).  PARSE_ERROR: not MessagePack.
inline ParseStatus parse(const void* p, size_t n, size_t& consumed, Value& out, size_t max_items = (size_t)1 << 20) {
  size_t pos = 0, budget = max_items;
  const ParseStatus st = detail::parse_at((const uint8_t*)p, n, pos, out, 0, budget);
  consumed = st == PARSE_OK ? pos : 0;
  return st;
}

}  // namespace mp

#endif
