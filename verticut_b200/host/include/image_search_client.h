// image_search_client - the reference's user-facing query interface (src/image_search_client.h:12-26:
// ping, search_image_by_id(id, knn, approximate) -> list<pair<image id, distance>>), served IN-PROCESS from the
// GPU-resident index instead of over msgpack-rpc to a server that forks `ssh ... mpirun` per query
// (src/image_search_server.cc:58-83).  Same method names, argument meaning and result order (descending
// distance, the order image_search_server::wait_and_parse collects the worker's "id : dist" lines in).
// Query by id needs the main table id -> code, which this implementation keeps; in the shipped
// reference that path is dead (src/distributed_image_search.cc:116).
// The msgpack-rpc transport itself is out of scope (library not available); a network front end would wrap
// exactly these two calls.
#ifndef VERTICUT_B200_IMAGE_SEARCH_CLIENT_H
#define VERTICUT_B200_IMAGE_SEARCH_CLIENT_H

#include <stdint.h>
#include <list>
#include <stdexcept>
#include <string>
#include <utility>

#include "gpu_search_worker.h"

class image_search_client {
 public:
  // in-process "connection": the proxy holding the tables (must outlive the client)
  explicit image_search_client(GpuTableProxy* proxy) : proxy_(proxy), worker_(proxy, (int)proxy->size()) {}

  // src/image_search_server.cc:52-54: the server echoes the content
  std::string ping(const std::string& content) { return content; }

  // nearest neighbours of the stored image `id`; throws std::runtime_error if the id is unknown
  // (the reference dies with "Can't find match", src/distributed_image_search.cc:98)
  std::list<std::pair<uint32_t, uint32_t> > search_image_by_id(uint32_t id, int knn, bool approximate = false) {
    ID key;
    BinaryCode code;
    key.set_id(id);
    if (proxy_->get(key, code) != PROXY_FOUND) throw std::runtime_error("Can't find match");
    std::list<SearchWorker::search_result_st> r = worker_.find(code.code().data(), code.code().size(), knn, approximate);
    std::list<std::pair<uint32_t, uint32_t> > out;
    for (std::list<SearchWorker::search_result_st>::iterator it = r.begin(); it != r.end(); ++it)
      out.push_back(std::make_pair(it->image_id, it->dist));
    return out;
  }

 protected:
  GpuTableProxy* proxy_;
  SearchWorker worker_;
};

#endif
