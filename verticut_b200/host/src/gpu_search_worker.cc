#include "gpu_search_worker.h"

#include <stdio.h>
#include <stdlib.h>

SearchWorker::SearchWorker(GpuTableProxy* proxy_clt, int image_total) : proxy_clt_(proxy_clt), image_total_(image_total) {}

namespace {
// canonical ascending (dist, id) -> the reference's output order, descending distance
std::list<SearchWorker::search_result_st> to_list(const uint32_t* ids, const uint32_t* dists, uint32_t count) {
  std::list<SearchWorker::search_result_st> out;
  for (uint32_t i = count; i-- > 0;) {
    SearchWorker::search_result_st r;
    r.image_id = ids[i];
    r.dist = dists[i];
    out.push_back(r);
  }
  return out;
}
}  // namespace

std::vector<std::list<SearchWorker::search_result_st> > SearchWorker::find_batch(const char* codes, size_t nbytes, size_t nq,
                                                                                 int knn, bool approximate) {
  std::vector<std::list<search_result_st> > out(nq);
  stats_.assign(nq, vc_query_stats());
  if ((int)nbytes != proxy_clt_->code_bytes() || knn <= 0) {
    fprintf(stderr, "SearchWorker: query of %zu bytes against %d-byte codes\n", nbytes, proxy_clt_->code_bytes());
    abort();                                    // the reference asserts (src/search_worker.cc:75)
  }
  if (nq == 0 || proxy_clt_->finalize() != 0) return out;
  std::vector<uint32_t> ids(nq * knn), dists(nq * knn), counts(nq);
  if (vc_search_mih(proxy_clt_->handle(), codes, (uint32_t)nq, (uint32_t)knn, approximate ? 1 : 0, -1, ids.data(), dists.data(),
                    counts.data(), stats_.data()) != VC_OK) {
    fprintf(stderr, "SearchWorker: %s\n", vc_last_error());
    abort();                                    // mpi_coordinator::die in the reference
  }
  for (size_t q = 0; q < nq; ++q) out[q] = to_list(&ids[q * knn], &dists[q * knn], counts[q]);
  return out;
}

std::list<SearchWorker::search_result_st> SearchWorker::find(const char* binary_code, size_t nbytes, int knn, bool approximate) {
  result_.clear();                              // per-query state is reset (src/search_worker.cc:68-73)
  std::vector<std::list<search_result_st> > r = find_batch(binary_code, nbytes, 1, knn, approximate);
  result_ = r[0];
  return result_;
}

std::list<SearchWorker::search_result_st> SearchWorker::linear_find(const char* binary_code, size_t nbytes, int knn) {
  std::list<search_result_st> out;
  if ((int)nbytes != proxy_clt_->code_bytes() || knn <= 0 || proxy_clt_->finalize() != 0) return out;
  std::vector<uint32_t> ids(knn), dists(knn);
  uint32_t count = 0;
  if (vc_search_linear(proxy_clt_->handle(), binary_code, 1, (uint32_t)knn, ids.data(), dists.data(), &count) != VC_OK) {
    fprintf(stderr, "SearchWorker: %s\n", vc_last_error());
    abort();
  }
  return to_list(ids.data(), dists.data(), count);
}

void SearchWorker::get_stat(uint64_t& n_main_reads, uint64_t& n_sub_reads, uint64_t& n_local_reads, uint32_t& radius) {
  n_main_reads = 0;
  n_sub_reads = stats_.empty() ? 0 : stats_[0].probes;
  n_local_reads = stats_.empty() ? 0 : stats_[0].occupancy_tests;
  radius = stats_.empty() ? 0 : stats_[0].radius;
}
