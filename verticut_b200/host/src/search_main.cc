// distributed-image-search / linear-search / accuracy-test / integrity-check for --server gpu, one binary:
//   image-search mih       -f codes -q queries [-b bits] [-n tables] [-k knn] [-a] [-i max images]
//   image-search linear    -f codes -q queries ...
//   image-search accuracy  -f codes -q queries ...      (reference: src/accuracy_test.cc)
//   image-search integrity -f codes ...                 (reference: src/integrity_check.cc)
//   image-search byid      -f codes -I <image id> ...   (reference: search_image_by_id, src/image_search_client.h:23-25;
//                                                        the query-by-id branch of src/distributed_image_search.cc:95-117)
// The reference's own command line is accepted as well (what src/run_distributed_search.py:74-79 and the RPC server's popen line,
// src/image_search_server.cc:58-67, start):
//   distributed-image-search <config> <image count> <binary bits> <substring bits> <k> <server> <read mode> <approximate> <query id> [query file]
// (src/distributed_image_search.cc:141-156).  <server> must be "gpu"; <config> is the GPU backend's server list, which names
// the device and - with an "index <file>" or "codes <file>" line - where the tables come from; with a query file the
// averaged statistics are printed as the reference does (:87-93), without one the neighbours of image <query id> ("id : dist").
// Output formats follow the reference: result lines "id : dist" (what image_search_server.cc:94 parses), the
// averaged statistics line of src/distributed_image_search.cc:87-93, and for the scan
// "Find image with id=%d and hamming_dist=%d" (src/linear_search.cc:62).  Neighbours are listed in descending
// distance, as the reference does.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include <iostream>
#include <vector>

#include "args_config.h"
#include "gpu_search_worker.h"
#include "image_search_client.h"
#include "integrity.h"

static double now() { timeval t; gettimeofday(&t, 0); return t.tv_sec + t.tv_usec * 1e-6; }

static std::vector<char> read_queries(const char* path, int rec, int limit) {
  std::vector<char> q;
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "Couldn't open file %s\n", path); exit(1); }
  std::vector<char> one(rec);
  while ((int)(q.size() / rec) < limit && fread(one.data(), rec, 1, f) == 1) q.insert(q.end(), one.begin(), one.end());
  fclose(f);
  return q;
}

static bool is_mode(const char* a) {
  return !strcmp(a, "mih") || !strcmp(a, "linear") || !strcmp(a, "accuracy") || !strcmp(a, "integrity") || !strcmp(a, "byid");
}

// src/distributed_image_search.cc:141-156 setup() + :36-127 main()
static int positional_main(int argc, char* argv[]) {
  if (argc < 10) { fprintf(stderr, "Incorrect number of arguments!\n"); return 1; }
  config_path = argv[1];
  image_total = atoi(argv[2]);
  binary_bits = atoi(argv[3]);
  const int n_local_bits = atoi(argv[4]);
  knn = atoi(argv[5]);
  read_mode = atoi(argv[7]);
  approximate = atoi(argv[8]);
  query_image_id = atoi(argv[9]);
  if (argc >= 11) query_file = argv[10];
  if (strcmp(argv[6], "gpu") != 0) { fprintf(stderr, "Unrecognized server type.\n"); return 1; }
  if (n_local_bits <= 0 || binary_bits % n_local_bits) { fprintf(stderr, "binary bits %d not divisible into %d-bit substrings\n", binary_bits, n_local_bits); return 1; }
  n_tables = binary_bits / n_local_bits;
  GpuTableProxy proxy(binary_bits, n_tables);
  double t0 = now();
  if (proxy.init(config_path) != 0) { fprintf(stderr, "proxy init failed: %s\n", proxy.last_error()); return 1; }
  if (!proxy.config_index_path().empty()) {
    if (proxy.load(proxy.config_index_path().c_str()) != 0) { fprintf(stderr, "Can't read index %s\n", proxy.config_index_path().c_str()); return 1; }
  } else if (!proxy.config_codes_path().empty()) {
    if (proxy.load_code_file(proxy.config_codes_path().c_str(), (uint64_t)image_total) != 0 || proxy.finalize() != 0) {
      fprintf(stderr, "Can't load %s\n", proxy.config_codes_path().c_str());
      return 1;
    }
  } else {
    fprintf(stderr, "%s names no tables: add a line \"index <file written by build-tables -o>\" or \"codes <raw code file>\"\n", config_path);
    return 1;
  }
  const double t_connect = now() - t0;
  const int rec = proxy.code_bytes();
  SearchWorker worker(&proxy, (int)proxy.size());
  double t1 = now();
  try {
    if (query_file) {
      std::vector<char> queries = read_queries(query_file, rec, max_queries);
      const size_t nq = queries.size() / rec;
      worker.find_batch(queries.data(), rec, nq, knn, approximate != 0);
      uint64_t sub = 0, local = 0, radius = 0;
      for (size_t q = 0; q < nq; ++q) { sub += worker.batch_stats()[q].probes; local += worker.batch_stats()[q].occupancy_tests; radius += worker.batch_stats()[q].radius; }
      if (nq) {
        std::cout << "Averate result : " << std::endl;     // sic, :88
        std::cout << 0 << "  n_main_reads : " << 0 << " , n_sub_reads : " << sub / nq << ", n_local_reads : " << local / nq
                  << ", radius : " << radius / nq << ", rdma : " << 0 << std::endl;
      }
    } else {
      // the branch that is dead in the shipped reference (:95-117, "The version without main table doesn't support query by id")
      image_search_client client(&proxy);
      std::list<std::pair<uint32_t, uint32_t> > r = client.search_image_by_id((uint32_t)query_image_id, knn, approximate != 0);
      for (std::list<std::pair<uint32_t, uint32_t> >::iterator it = r.begin(); it != r.end(); ++it)
        std::cout << it->first << " : " << it->second << std::endl;
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  const double t_while = now() - t1;
  std::cout << "-------Timings-------" << std::endl;      // src/timer.h:31-36
  std::cout << "connect : " << t_connect << " s" << std::endl;
  std::cout << "while : " << t_while << " s" << std::endl;
  std::cout << "---------------------" << std::endl;
  proxy.close();
  return 0;
}

int main(int argc, char* argv[]) {
  if (argc < 2) usage();
  if (!is_mode(argv[1]) && argv[1][0] != '-') return positional_main(argc, argv);
  const char* mode = argv[1];
  configure(argc - 1, argv + 1);
  int rec = binary_bits / 8;
  const bool linear = !strcmp(mode, "linear");
  GpuTableProxy proxy(binary_bits, linear ? 0 : n_tables);
  double t0 = now();
  if (proxy.init(config_path) != 0) { fprintf(stderr, "proxy init failed: %s\n", proxy.last_error()); return 1; }
  if (index_in) {
    if (proxy.load(index_in) != 0) { fprintf(stderr, "Can't read index %s\n", index_in); return 1; }
  } else {
    if (proxy.load_code_file(binary_file, (uint64_t)image_total) != 0) { fprintf(stderr, "Can't open file %s.\n", binary_file); return 1; }
    if (proxy.finalize() != 0) { fprintf(stderr, "build failed: %s\n", proxy.last_error()); return 1; }
  }
  rec = proxy.code_bytes();
  double t_connect = now() - t0;
  if (!strcmp(mode, "integrity")) return run_integrity(&proxy);
  if (!strcmp(mode, "byid")) {
    if (query_image_id < 0) { fprintf(stderr, "an image id is required (-I)\n"); return 1; }
    image_search_client client(&proxy);
    try {
      std::list<std::pair<uint32_t, uint32_t> > r = client.search_image_by_id((uint32_t)query_image_id, knn, approximate != 0);
      for (std::list<std::pair<uint32_t, uint32_t> >::iterator it = r.begin(); it != r.end(); ++it)
        std::cout << it->first << " : " << it->second << std::endl;      // the format image_search_server.cc:94 parses
    } catch (const std::exception& e) {
      fprintf(stderr, "%s\n", e.what());
      return 1;
    }
    proxy.close();
    return 0;
  }
  if (!query_file) { fprintf(stderr, "a query file is required (-q)\n"); return 1; }
  std::vector<char> queries = read_queries(query_file, rec, max_queries);
  const size_t nq = queries.size() / rec;
  SearchWorker worker(&proxy, (int)proxy.size());
  typedef std::list<SearchWorker::search_result_st> Result;
  double t1 = now();
  if (linear) {
    for (size_t q = 0; q < nq; ++q) {
      Result r = worker.linear_find(&queries[q * rec], rec, knn);
      for (Result::iterator it = r.begin(); it != r.end(); ++it)
        printf("Find image with id=%d and hamming_dist=%d\n", it->image_id, it->dist);
    }
  } else if (!strcmp(mode, "mih")) {
    std::vector<Result> res = worker.find_batch(queries.data(), rec, nq, knn, approximate != 0);
    uint64_t sub = 0, local = 0, radius = 0;
    for (size_t q = 0; q < nq; ++q) {
      printf("query %zu\n", q);
      for (Result::iterator it = res[q].begin(); it != res[q].end(); ++it) std::cout << it->image_id << " : " << it->dist << std::endl;
      sub += worker.batch_stats()[q].probes; local += worker.batch_stats()[q].occupancy_tests; radius += worker.batch_stats()[q].radius;
    }
    if (nq) {
      std::cout << "Averate result : " << std::endl;     // sic, src/distributed_image_search.cc:88
      std::cout << 0 << "  n_main_reads : " << 0 << " , n_sub_reads : " << sub / nq << ", n_local_reads : " << local / nq
                << ", radius : " << radius / nq << ", rdma : " << 0 << std::endl;
    }
  } else if (!strcmp(mode, "accuracy")) {
    // approximate vs exact: mean distance of both and the count of approximate results that are worse than
    // the exact farthest one (src/accuracy_test.cc:106-135)
    uint64_t total_ex = 0, total_app = 0, inaccurate = 0;
    std::vector<Result> app = worker.find_batch(queries.data(), rec, nq, knn, true);
    std::vector<Result> ex = worker.find_batch(queries.data(), rec, nq, knn, false);
    for (size_t q = 0; q < nq; ++q) {
      for (Result::iterator it = ex[q].begin(); it != ex[q].end(); ++it) total_ex += it->dist;
      for (Result::iterator it = app[q].begin(); it != app[q].end(); ++it) total_app += it->dist;
      const uint32_t threshold = ex[q].empty() ? 0 : ex[q].front().dist;
      for (Result::iterator it = app[q].begin(); it != app[q].end() && it->dist > threshold; ++it) ++inaccurate;
      std::cout << (float)total_ex / (q + 1) / knn << " " << (float)total_app / (q + 1) / knn << " " << (float)inaccurate / (q + 1) / knn << std::endl;
    }
  } else {
    usage();
  }
  double t_while = now() - t1;
  std::cout << "-------Timings-------" << std::endl;      // src/timer.h:31-36
  std::cout << "connect : " << t_connect << " s" << std::endl;
  std::cout << "while : " << t_while << " s" << std::endl;
  std::cout << "---------------------" << std::endl;
  proxy.close();
  return 0;
}
