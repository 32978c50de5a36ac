// Table integrity check of the reference (src/integrity_check.cc:25-35,52-67): re-read the code file and
// require every code to be present, with its id, in its bucket of every table - through BaseProxy::get.
#ifndef VERTICUT_B200_INTEGRITY_H
#define VERTICUT_B200_INTEGRITY_H
#include <stdio.h>
#include <string.h>
#include <vector>
#include "args_config.h"
#include "gpu_table_proxy.h"

static inline int run_integrity(GpuTableProxy* proxy) {
  // every code must be found, with its id, in its bucket of every table (src/integrity_check.cc:25-35,52-67)
  FILE* fh = fopen(binary_file, "rb");
  if (!fh) { fprintf(stderr, "Can't open file %s.", binary_file); return 1; }
  const int rec = binary_bits / 8, substr_len = rec / n_tables;
  std::vector<char> code(rec);
  Image_List img_list;
  HashIndex idx;
  uint32_t id = 0, bad = 0;
  while ((int)id < image_total && fread(code.data(), rec, 1, fh) == 1) {
    for (int t = 0; t < n_tables; ++t) {
      uint32_t index = 0;
      for (int i = substr_len - 1; i >= 0; --i) index = (index << 8) | (uint8_t)code[t * substr_len + i];
      idx.set_table_id(t);
      idx.set_index(index);
      bool ok = false;
      if (proxy->get(idx, img_list) == PROXY_FOUND)
        for (int i = 0; i < img_list.images_size() && !ok; ++i)
          ok = img_list.images(i).id() == id && memcmp(img_list.images(i).code().data(), code.data(), rec) == 0;
      if (!ok) { if (bad++ < 10) fprintf(stderr, "integrity: image %u missing from table %d bucket %u\n", id, t, index); }
    }
    ++id;
  }
  fclose(fh);
  printf("integrity check : %u images, %u errors\n", id, bad);
  return bad ? 1 : 0;
}


#endif
