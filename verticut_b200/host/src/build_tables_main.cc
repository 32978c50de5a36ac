// build-tables (reference: src/build_hash_tables.cc).  Reads the raw code file and builds the n_tables
// substring tables - here in the HBM of the selected GPU.  With --via-put the tool drives the proxy exactly
// like the reference's load_binarycode loop (get -> append -> put per code and table), to exercise the
// BaseProxy compatibility path; the default streams the file straight to the device.
// --bitmaps also writes the per-table bucket-occupancy bitmaps in the format of the reference's generate_bitmap tool
// (src/generate_bitmap.cc:84-87,119-125: one raw file of 2^s bits per table, "<code file>_bmp_<table 1..m>_2b_4k.raw",
// bit i of 32-bit word i / 32 = bucket i is non-empty; 512 MiB per table for s = 32) - what bitmap_deamon.cc:41-61 maps.
#include <stdio.h>
#include <string.h>
#include <sys/time.h>

#include <vector>

#include "args_config.h"
#include "gpu_table_proxy.h"
#include "integrity.h"

static double now() { timeval t; gettimeofday(&t, 0); return t.tv_sec + t.tv_usec * 1e-6; }

static int load_via_put(GpuTableProxy* proxy, const char* fname) {
  FILE* fh = fopen(fname, "rb");
  if (!fh) { fprintf(stderr, "Can't open file %s.", fname); return -1; }
  const int rec = binary_bits / 8, substr_len = rec / n_tables;
  std::vector<char> code(rec);
  Image_List img_list;
  HashIndex idx;
  uint32_t id = 0;
  while ((int)id < image_total && fread(code.data(), rec, 1, fh) == 1) {
    for (int t = 0; t < n_tables; ++t) {          // the reference runs one MPI rank per table
      uint32_t index = 0;
      for (int i = substr_len - 1; i >= 0; --i) index = (index << 8) | (uint8_t)code[t * substr_len + i];
      idx.set_table_id(t);
      idx.set_index(index);
      if (proxy->get(idx, img_list) != PROXY_FOUND) img_list.clear_images();
      ID_Code_Pair* pair = img_list.add_images();
      pair->set_id(id);
      pair->set_code(code.data(), rec);
      if (proxy->put(idx, img_list) != PROXY_PUT_DONE) { fclose(fh); return -1; }
    }
    ++id;
  }
  fclose(fh);
  return 0;
}

int main(int argc, char* argv[]) {
  bool via_put = false, check = false, bitmaps = false;
  std::vector<char*> args;
  for (int i = 0; i < argc; ++i) {
    if (!strcmp(argv[i], "--via-put")) via_put = true;
    else if (!strcmp(argv[i], "--check")) check = true;
    else if (!strcmp(argv[i], "--bitmaps")) bitmaps = true;
    else args.push_back(argv[i]);
  }
  configure((int)args.size(), args.data());
  GpuTableProxy proxy(binary_bits, n_tables);
  if (proxy.init(config_path) != 0) { fprintf(stderr, "proxy init failed: %s\n", proxy.last_error()); return 1; }
  double t0 = now();
  int rc = via_put ? load_via_put(&proxy, binary_file) : proxy.load_code_file(binary_file, (uint64_t)image_total);
  if (rc != 0) { fprintf(stderr, "Can't load %s\n", binary_file); return 1; }
  double t1 = now();
  if (proxy.finalize() != 0) { fprintf(stderr, "build failed: %s\n", proxy.last_error()); return 1; }
  double t2 = now();
  vc_index_info info;
  vc_index_get_info(proxy.handle(), &info);
  printf("images : %llu, tables : %u x %u-bit, load : %.3f s, build : %.3f s, device bytes : %llu\n",
         (unsigned long long)info.n_codes, info.n_tables, info.substring_bits, t1 - t0, t2 - t1, (unsigned long long)info.device_bytes);
  if (index_out && proxy.save(index_out) != 0) { fprintf(stderr, "can't write %s: %s\n", index_out, proxy.last_error()); return 1; }
  if (bitmaps) {
    const uint64_t n_words = (1ull << info.substring_bits) / 32;
    std::vector<uint32_t> words(n_words);
    for (uint32_t t = 0; t < info.n_tables; ++t) {
      if (vc_occupancy_bitmap_get(proxy.handle(), t, words.data(), n_words) != VC_OK) { fprintf(stderr, "bitmap of table %u: %s\n", t, proxy.last_error()); return 1; }
      char name[4096];
      snprintf(name, sizeof name, "%s_bmp_%u_2b_4k.raw", binary_file, t + 1);
      FILE* bf = fopen(name, "wb");
      if (!bf || fwrite(words.data(), 4, n_words, bf) != n_words) { fprintf(stderr, "can't create files.\n"); return 1; }
      fclose(bf);
    }
  }
  int rc2 = check ? run_integrity(&proxy) : 0;
  proxy.close();
  return rc2;
}
