// SearchWorker - the reference's query class (src/search_worker.h:18-64) over the GPU-resident tables.
// Same public surface: search_result_st, find(code, nbytes, knn, approximate) returning the neighbours in
// DESCENDING distance (the reference pops a max-heap into the list, src/search_worker.cc:210-216), get_knn(),
// get_stat().  The mpi_coordinator argument of the reference constructor is gone: there are no per-table
// ranks to coordinate, the tables of all m substrings live in one GPU (or one id-shard each, see
// verticut_b200/sharded.py).  find_batch() is the form the GPU wants: many queries per call.
#ifndef VERTICUT_B200_GPU_SEARCH_WORKER_H
#define VERTICUT_B200_GPU_SEARCH_WORKER_H

#include <stdint.h>
#include <list>
#include <vector>

#include "gpu_table_proxy.h"

#define APPROXIMATE_FACTOR 20   // src/search_worker.h:14

class SearchWorker {
 public:
  struct search_result_st {
    uint32_t image_id;
    uint32_t dist;
  };

  SearchWorker(GpuTableProxy* proxy_clt, int image_total);

  // one query; nbytes must equal the index's code size (the reference asserts divisibility by the table count)
  std::list<search_result_st> find(const char* binary_code, size_t nbytes, int knn, bool approximate);
  // nq queries back to back in `codes`; results[q] as find() would return them
  std::vector<std::list<search_result_st> > find_batch(const char* codes, size_t nbytes, size_t nq, int knn, bool approximate);
  // brute-force scan of the same index (search_K_nearest_neighbors of src/linear_search.cc:39-64), descending distance
  std::list<search_result_st> linear_find(const char* binary_code, size_t nbytes, int knn);

  std::list<search_result_st> get_knn() { return result_; }
  // statistics of the last find(): n_sub_reads = bucket probes issued, n_local_reads = occupancy-bitmap tests,
  // n_main_reads = 0 (never incremented by the reference either), radius = last radius searched
  void get_stat(uint64_t& n_main_reads, uint64_t& n_sub_reads, uint64_t& n_local_reads, uint32_t& radius);
  const std::vector<vc_query_stats>& batch_stats() const { return stats_; }

 protected:
  GpuTableProxy* proxy_clt_;
  std::list<search_result_st> result_;
  std::vector<vc_query_stats> stats_;
  int image_total_;
};

#endif
