// pb-wire - command-line access to the hand-written KV-schema messages (image_search.pb.h) for the wire-format test
// against the real protobuf implementation (tests/test_pb_wire.py).  One command per stdin line, one answer per line:
//   enc ID <id> | enc BinaryCode <hex> | enc HashIndex <table> <index> | enc ID_Code_Pair <id> <hex>
//   enc Image_List <id>:<hex> ... | enc ImageList <u32> ...                      -> the serialised bytes in hex ("-" = empty)
//   dec <Type> <hex>   -> the fields in the same notation as the enc arguments, or "ERROR" if the bytes do not parse
#include <stdio.h>
#include <stdlib.h>

#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "image_search.pb.h"

static std::string to_hex(const std::string& b) {
  if (b.empty()) return "-";
  static const char* d = "0123456789abcdef";
  std::string h;
  for (size_t i = 0; i < b.size(); ++i) { h.push_back(d[(unsigned char)b[i] >> 4]); h.push_back(d[(unsigned char)b[i] & 15]); }
  return h;
}
static std::string from_hex(const std::string& h) {
  std::string b;
  if (h == "-") return b;
  for (size_t i = 0; i + 1 < h.size(); i += 2) b.push_back((char)strtoul(h.substr(i, 2).c_str(), 0, 16));
  return b;
}

int main() {
  std::string line;
  while (std::getline(std::cin, line)) {
    std::istringstream in(line);
    std::string cmd, type;
    in >> cmd >> type;
    std::vector<std::string> a;
    for (std::string t; in >> t;) a.push_back(t);
    std::string out;
    if (cmd == "enc") {
      if (type == "ID") { ID m; m.set_id((uint32_t)strtoul(a[0].c_str(), 0, 10)); m.SerializeToString(&out); }
      else if (type == "BinaryCode") { BinaryCode m; m.set_code(from_hex(a[0])); m.SerializeToString(&out); }
      else if (type == "HashIndex") { HashIndex m; m.set_table_id((uint32_t)strtoul(a[0].c_str(), 0, 10)); m.set_index((uint32_t)strtoul(a[1].c_str(), 0, 10)); m.SerializeToString(&out); }
      else if (type == "ID_Code_Pair") { ID_Code_Pair m; m.set_id((uint32_t)strtoul(a[0].c_str(), 0, 10)); m.set_code(from_hex(a[1])); m.SerializeToString(&out); }
      else if (type == "Image_List") {
        Image_List m;
        for (size_t i = 0; i < a.size(); ++i) {
          const size_t c = a[i].find(':');
          ID_Code_Pair* p = m.add_images();
          p->set_id((uint32_t)strtoul(a[i].substr(0, c).c_str(), 0, 10)); p->set_code(from_hex(a[i].substr(c + 1)));
        }
        m.SerializeToString(&out);
      } else if (type == "ImageList") { ImageList m; for (size_t i = 0; i < a.size(); ++i) m.add_images((uint32_t)strtoul(a[i].c_str(), 0, 10)); m.SerializeToString(&out); }
      std::cout << to_hex(out) << std::endl;
    } else if (cmd == "dec") {
      const std::string bytes = from_hex(a.empty() ? "-" : a[0]);
      std::ostringstream o;
      bool ok = false;
      if (type == "ID") { ID m; ok = m.ParseFromString(bytes); o << m.id(); }
      else if (type == "BinaryCode") { BinaryCode m; ok = m.ParseFromString(bytes); o << to_hex(m.code()); }
      else if (type == "HashIndex") { HashIndex m; ok = m.ParseFromString(bytes); o << m.table_id() << " " << m.index(); }
      else if (type == "ID_Code_Pair") { ID_Code_Pair m; ok = m.ParseFromString(bytes); o << m.id() << " " << to_hex(m.code()); }
      else if (type == "Image_List") { Image_List m; ok = m.ParseFromString(bytes); for (int i = 0; i < m.images_size(); ++i) o << (i ? " " : "") << m.images(i).id() << ":" << to_hex(m.images(i).code()); }
      else if (type == "ImageList") { ImageList m; ok = m.ParseFromString(bytes); for (int i = 0; i < m.images_size(); ++i) o << (i ? " " : "") << m.images(i); }
      std::cout << (ok ? o.str() : std::string("ERROR")) << std::endl;
    }
  }
  return 0;
}
