// image_search_client - the reference's user-facing query interface (src/image_search_client.h:12-26:
// ping, search_image_by_id(id, knn, approximate) -> list<pair<image id, distance>>), served IN-PROCESS from the
// GPU-resident index instead of over msgpack-rpc to a server that forks `ssh ... mpirun` per query
// (src/image_search_server.cc:58-83).  Same method names, argument meaning and result order (descending
// distance, the order image_search_server::wait_and_parse collects the worker's "id : dist" lines in).
// Query by id needs the main table id -> code, which this implementation keeps; in the shipped
// reference that path is dead (src/distributed_image_search.cc:116).
// Two kinds of connection: in-process (constructed from the proxy that holds the tables), or - the reference's own
// constructor, image_search_client(ip, port) - msgpack-rpc over TCP to an image-search-server (image_search_rpc.h,
// host/src/rpc_server_main.cc), whose wire format is the reference's.
#ifndef VERTICUT_B200_IMAGE_SEARCH_CLIENT_H
#define VERTICUT_B200_IMAGE_SEARCH_CLIENT_H

#include <stdint.h>
#include <list>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>

#include "gpu_search_worker.h"
#include "image_search_rpc.h"

class image_search_client {
 public:
  // in-process "connection": the proxy holding the tables (must outlive the client)
  explicit image_search_client(GpuTableProxy* proxy) : proxy_(proxy), worker_(new SearchWorker(proxy, (int)proxy->size())) {}
  // src/image_search_client.cc:3-9: a server's address; connects on the first call
  image_search_client(std::string ip, uint16_t port) : proxy_(0), remote_(new vcrpc::remote_client(ip, port)) {}

  // src/image_search_server.cc:52-54: the server echoes the content
  std::string ping(const std::string& content) { return remote_ ? remote_->ping(content) : content; }

  // nearest neighbours of the stored image `id`; throws std::runtime_error if the id is unknown
  // (the reference dies with "Can't find match", src/distributed_image_search.cc:98) - over RPC the server sends
  // that message back as the call's error (req.error(e.what()), src/image_search_server.cc:46) and it is thrown here
  std::list<std::pair<uint32_t, uint32_t> > search_image_by_id(uint32_t id, int knn, bool approximate = false) {
    if (remote_) return remote_->search_image_by_id(id, knn, approximate);
    ID key;
    BinaryCode code;
    key.set_id(id);
    if (proxy_->get(key, code) != PROXY_FOUND) throw std::runtime_error("Can't find match");
    std::list<SearchWorker::search_result_st> r = worker_->find(code.code().data(), code.code().size(), knn, approximate);
    std::list<std::pair<uint32_t, uint32_t> > out;
    for (std::list<SearchWorker::search_result_st>::iterator it = r.begin(); it != r.end(); ++it)
      out.push_back(std::make_pair(it->image_id, it->dist));
    return out;
  }

 protected:
  GpuTableProxy* proxy_;
  std::unique_ptr<SearchWorker> worker_;
  std::unique_ptr<vcrpc::remote_client> remote_;
};

#endif
