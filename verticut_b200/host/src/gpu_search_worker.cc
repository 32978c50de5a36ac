#include "gpu_search_worker.h"

#include <stdio.h>
#include <stdlib.h>

#include <stdexcept>
#include <string>

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
  // The reference asserts on a malformed query (src/search_worker.cc:75) - there that kills one forked worker, here the
  // process holds the only copy of the GPU index, so argument errors are thrown to the caller instead (the RPC layer
  // turns them into the call's error string, image_search_rpc.h handle_message).
  if ((int)nbytes != proxy_clt_->code_bytes())
    throw std::invalid_argument("SearchWorker: query of " + std::to_string(nbytes) + " bytes against " +
                                std::to_string(proxy_clt_->code_bytes()) + "-byte codes");
  if (knn <= 0 || knn > (int)VC_MAX_K)
    throw std::invalid_argument("SearchWorker: knn must be in [1, " + std::to_string(VC_MAX_K) + "]");
  if (nq == 0 || proxy_clt_->finalize() != 0) return out;
  std::vector<uint32_t> ids(nq * knn), dists(nq * knn), counts(nq);
  if (vc_search_mih(proxy_clt_->handle(), codes, (uint32_t)nq, (uint32_t)knn, approximate ? 1 : 0, -1, ids.data(), dists.data(),
                    counts.data(), stats_.data()) != VC_OK)
    throw std::runtime_error(std::string("SearchWorker: ") + vc_last_error());      // mpi_coordinator::die in the reference
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
  if ((int)nbytes != proxy_clt_->code_bytes() || knn <= 0 || knn > (int)VC_MAX_K || proxy_clt_->finalize() != 0) return out;
  std::vector<uint32_t> ids(knn), dists(knn);
  uint32_t count = 0;
  if (vc_search_linear(proxy_clt_->handle(), binary_code, 1, (uint32_t)knn, ids.data(), dists.data(), &count) != VC_OK)
    throw std::runtime_error(std::string("SearchWorker: ") + vc_last_error());
  return to_list(ids.data(), dists.data(), count);
}

void SearchWorker::get_stat(uint64_t& n_main_reads, uint64_t& n_sub_reads, uint64_t& n_local_reads, uint32_t& radius) {
  n_main_reads = 0;
  n_sub_reads = stats_.empty() ? 0 : stats_[0].probes;
  n_local_reads = stats_.empty() ? 0 : stats_[0].occupancy_tests;
  radius = stats_.empty() ? 0 : stats_[0].radius;
}
