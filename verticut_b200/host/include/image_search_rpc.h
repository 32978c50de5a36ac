// The reference's query RPC on the wire: msgpack-rpc over TCP with the two methods of
// src/image_search_server.cc:22-50 -
//     "ping"(string) -> string                                      (:52-54, echoes the content)
//     "search_image_by_id"(uint32 id, uint32 knn, bool approximate) -> array of [image id, distance]
//                                                                   (:58-83; list<pair<uint32,uint32>>)
// and its error behaviour: unknown method -> NO_METHOD_ERROR, parameters that do not convert -> ARGUMENT_ERROR,
// any other exception -> its what() string (:37-48).  msgpack-rpc messages are plain MessagePack arrays sent back
// to back:  request [0, msgid, method, params]   response [1, msgid, error, result]   notification [2, method, params].
//
// The reference's server forks `ssh worker ... mpirun` per query and parses the worker's "id : dist" lines
// (:58-100); this one answers from a `service` object in the same process (the GPU-resident index behind the
// in-process image_search_client, see rpc_server_main.cc).  msgpack-rpc / mpio are not available here, so the
// transport is hand-written on POSIX sockets: worker threads share one listening socket (the reference runs
// `instance.run(n_threads)` with 10 threads, src/image_server_main.cc:13,90), one connection per thread at a time.
//
// Header-only; depends on nothing but msgpack_lite.h and the C library, so the wire format is testable without a GPU.
#ifndef VERTICUT_B200_IMAGE_SEARCH_RPC_H
#define VERTICUT_B200_IMAGE_SEARCH_RPC_H

#include <arpa/inet.h>
#include <errno.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/types.h>
#include <unistd.h>

#include <atomic>
#include <list>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "msgpack_lite.h"

namespace vcrpc {

typedef std::list<std::pair<uint32_t, uint32_t> > result_list;   // (image id, distance), in the order the server produced them

enum { REQUEST = 0, RESPONSE = 1, NOTIFY = 2 };                  // msgpack-rpc message types
enum { NO_METHOD_ERROR = 1, ARGUMENT_ERROR = 2 };                // msgpack-rpc's built-in error codes

// what a server answers from
struct service {
  virtual ~service() {}
  virtual std::string ping(const std::string& content) { return content; }
  virtual result_list search_image_by_id(uint32_t id, uint32_t knn, bool approximate) = 0;
};

struct rpc_error : std::runtime_error {
  int code;   // NO_METHOD_ERROR / ARGUMENT_ERROR, 0 when the server sent a message string (or the transport failed)
  rpc_error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// ---- messages ----------------------------------------------------------------------------------------------------
inline void pack_result_list(mp::Packer& pk, const result_list& r) {
  pk.pack_array(r.size());
  for (result_list::const_iterator it = r.begin(); it != r.end(); ++it) { pk.pack_array(2); pk.pack_uint(it->first); pk.pack_uint(it->second); }
}
inline bool unpack_result_list(const mp::Value& v, result_list& out) {
  if (!v.is_array()) return false;
  out.clear();
  for (size_t i = 0; i < v.a.size(); ++i) {
    const mp::Value& e = v.a[i];
    if (!e.is_array() || e.a.size() < 2 || !e.a[0].is_uint() || !e.a[1].is_uint() || e.a[0].u > 0xffffffffull || e.a[1].u > 0xffffffffull) return false;
    out.push_back(std::make_pair((uint32_t)e.a[0].u, (uint32_t)e.a[1].u));
  }
  return true;
}
inline std::string request_ping(uint32_t msgid, const std::string& content) {
  mp::Packer pk;
  pk.pack_array(4); pk.pack_uint(REQUEST); pk.pack_uint(msgid); pk.pack_str("ping", 4);
  pk.pack_array(1); pk.pack_str(content);
  return pk.bytes();
}
inline std::string request_search(uint32_t msgid, uint32_t id, uint32_t knn, bool approximate) {
  mp::Packer pk;
  pk.pack_array(4); pk.pack_uint(REQUEST); pk.pack_uint(msgid); pk.pack_str("search_image_by_id", 18);
  pk.pack_array(3); pk.pack_uint(id); pk.pack_uint(knn); pk.pack_bool(approximate);
  return pk.bytes();
}

// One decoded message in, the bytes of the answer out ("" when none is due: notifications, responses, garbage).
// Mirrors image_search_server::dispatch (src/image_search_server.cc:22-50).
inline std::string handle_message(service& svc, const mp::Value& msg) {
  if (!msg.is_array() || msg.a.empty() || !msg.a[0].is_uint()) return std::string();
  const uint64_t type = msg.a[0].u;
  const bool is_request = type == REQUEST && msg.a.size() == 4 && msg.a[1].is_uint();
  const bool is_notify = type == NOTIFY && msg.a.size() == 3;
  if (!is_request && !is_notify) return std::string();
  const mp::Value& method = msg.a[is_request ? 2 : 1];
  const mp::Value& params = msg.a[is_request ? 3 : 2];
  mp::Packer pk;
  pk.pack_array(4); pk.pack_uint(RESPONSE); pk.pack_uint(is_request ? msg.a[1].u : 0);
  const size_t head = pk.bytes().size();
  try {
    if (method.is_str() && method.s == "ping") {
      // msgpack::type::tuple<std::string>: an array with at least one element, the first one raw
      if (!params.is_array() || params.a.size() < 1 || !params.a[0].is_str()) { pk.pack_uint(ARGUMENT_ERROR); pk.pack_nil(); }
      else { const std::string r = svc.ping(params.a[0].s); pk.pack_nil(); pk.pack_str(r); }
    } else if (method.is_str() && method.s == "search_image_by_id") {
      // msgpack::type::tuple<uint32_t, uint32_t, bool>
      if (!params.is_array() || params.a.size() < 3 || !params.a[0].is_uint() || params.a[0].u > 0xffffffffull ||
          !params.a[1].is_uint() || params.a[1].u > 0xffffffffull || params.a[2].type != mp::Value::BOOL) {
        pk.pack_uint(ARGUMENT_ERROR); pk.pack_nil();
      } else {
        const result_list r = svc.search_image_by_id((uint32_t)params.a[0].u, (uint32_t)params.a[1].u, params.a[2].b);
        pk.pack_nil(); pack_result_list(pk, r);
      }
    } else {
      pk.pack_uint(NO_METHOD_ERROR); pk.pack_nil();
    }
  } catch (const std::exception& e) {
    pk.bytes().resize(head);                                  // req.error(std::string(e.what()))
    pk.pack_str(std::string(e.what())); pk.pack_nil();
  }
  return is_request ? pk.bytes() : std::string();
}

// ---- sockets -----------------------------------------------------------------------------------------------------
namespace detail {
inline bool send_all(int fd, const std::string& s) {
  size_t off = 0;
  while (off < s.size()) {
    const ssize_t w = ::send(fd, s.data() + off, s.size() - off, MSG_NOSIGNAL);
    if (w < 0) { if (errno == EINTR) continue; return false; }
    off += (size_t)w;
  }
  return true;
}
// appends what arrives to buf until one whole MessagePack object sits at its front; 1 = got one, 0 = peer closed, -1 = error
// max_bytes / max_items bound what one message may cost: the buffer is re-parsed after every recv(), which is O(max_items)
// per attempt (strings are length-checked before they are copied), and the decoder never allocates more than max_items values.
inline int recv_object(int fd, std::string& buf, mp::Value& out, size_t max_bytes, size_t max_items = (size_t)1 << 20) {
  for (;;) {
    if (!buf.empty()) {
      size_t used = 0;
      const mp::ParseStatus st = mp::parse(buf.data(), buf.size(), used, out, max_items);
      if (st == mp::PARSE_OK) { buf.erase(0, used); return 1; }
      if (st == mp::PARSE_ERROR || buf.size() > max_bytes) return -1;
    }
    char tmp[65536];
    const ssize_t r = ::recv(fd, tmp, sizeof tmp, 0);
    if (r == 0) return buf.empty() ? 0 : -1;
    if (r < 0) { if (errno == EINTR) continue; return -1; }
    buf.append(tmp, (size_t)r);
  }
}
}  // namespace detail

// ---- server ------------------------------------------------------------------------------------------------------
class image_search_server {
 public:
  explicit image_search_server(service* svc) : svc_(svc), lfd_(-1), port_(0), stop_(false) {}
  ~image_search_server() { stop(); }

  // binds and listens; port 0 picks a free port.  Returns the bound port, or -1 (errno says why).
  int listen(const std::string& ip, uint16_t port) {
    lfd_ = ::socket(AF_INET, SOCK_STREAM, 0);
    if (lfd_ < 0) return -1;
    int one = 1;
    ::setsockopt(lfd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in a;
    memset(&a, 0, sizeof a);
    a.sin_family = AF_INET; a.sin_port = htons(port);
    if (::inet_pton(AF_INET, ip.c_str(), &a.sin_addr) != 1) { ::close(lfd_); lfd_ = -1; errno = EINVAL; return -1; }
    if (::bind(lfd_, (sockaddr*)&a, sizeof a) != 0 || ::listen(lfd_, 64) != 0) { const int e = errno; ::close(lfd_); lfd_ = -1; errno = e; return -1; }
    socklen_t len = sizeof a;
    ::getsockname(lfd_, (sockaddr*)&a, &len);
    port_ = ntohs(a.sin_port);
    return port_;
  }
  int port() const { return port_; }

  // n_threads workers accept and serve connections until stop(); start() returns at once, run() blocks
  void start(int n_threads) {
    for (int i = 0; i < (n_threads < 1 ? 1 : n_threads); ++i) workers_.push_back(std::thread(&image_search_server::worker, this));
  }
  void run(int n_threads) { start(n_threads); join(); }
  void join() { for (size_t i = 0; i < workers_.size(); ++i) if (workers_[i].joinable()) workers_[i].join(); workers_.clear(); }
  void stop() {
    stop_ = true;
    if (lfd_ >= 0) ::shutdown(lfd_, SHUT_RDWR);
    { std::lock_guard<std::mutex> g(mu_); for (std::set<int>::iterator it = conns_.begin(); it != conns_.end(); ++it) ::shutdown(*it, SHUT_RDWR); }
    join();
    if (lfd_ >= 0) { ::close(lfd_); lfd_ = -1; }
  }

 private:
  void worker() {
    while (!stop_) {
      const int fd = ::accept(lfd_, 0, 0);
      if (fd < 0) {
        if (stop_ || errno == EBADF || errno == EINVAL || errno == ENOTSOCK) break;      // the listening socket is gone
        if (errno != EINTR && errno != ECONNABORTED) usleep(10000);                       // out of descriptors / memory: try again
        continue;
      }
      int one = 1;
      ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
      { std::lock_guard<std::mutex> g(mu_); conns_.insert(fd); }
      std::string buf;
      mp::Value msg;
      while (!stop_ && detail::recv_object(fd, buf, msg, kMaxMessage, kMaxItems) == 1) {
        const std::string reply = handle_message(*svc_, msg);
        if (!reply.empty() && !detail::send_all(fd, reply)) break;
      }
      { std::lock_guard<std::mutex> g(mu_); conns_.erase(fd); }
      ::close(fd);
    }
  }
  static const size_t kMaxMessage = 1u << 20;   // requests are tens of bytes (a ping may carry a long string); a peer that streams more without completing an object is dropped
  static const size_t kMaxItems = 16;           // [type, msgid, method, [id, knn, approximate]] = 7 items: the decoder never allocates more, and an
                                                // incomplete message costs O(1) to re-parse after each recv() (a string's length prefix is checked first)
  service* svc_;
  int lfd_;
  int port_;
  std::atomic<bool> stop_;
  std::mutex mu_;
  std::set<int> conns_;
  std::vector<std::thread> workers_;
};

// ---- client ------------------------------------------------------------------------------------------------------
// One TCP connection, synchronous calls (the reference's client blocks on future::get as well, src/image_search_client.cc:19-34).
class remote_client {
 public:
  remote_client(const std::string& ip, uint16_t port, int timeout_s = 480 /* s.set_timeout(120 * 4), :29 */)
      : ip_(ip), port_(port), timeout_s_(timeout_s), fd_(-1), msgid_(0) {}
  ~remote_client() { disconnect(); }

  std::string ping(const std::string& content) {
    const mp::Value r = call(request_ping(next_id(), content));
    if (!r.is_str()) throw rpc_error(0, "ping: the result is not a string");
    return r.s;
  }
  result_list search_image_by_id(uint32_t id, int knn, bool approximate = false) {
    const mp::Value r = call(request_search(next_id(), id, (uint32_t)knn, approximate));
    result_list out;
    if (!unpack_result_list(r, out)) throw rpc_error(0, "search_image_by_id: the result is not a list of (id, distance) pairs");
    return out;
  }

 private:
  uint32_t next_id() { return msgid_++; }
  void disconnect() { if (fd_ >= 0) { ::close(fd_); fd_ = -1; } buf_.clear(); }
  void ensure_connected() {
    if (fd_ >= 0) return;
    addrinfo hints, *res = 0;
    memset(&hints, 0, sizeof hints);
    hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
    char portstr[16];
    snprintf(portstr, sizeof portstr, "%u", (unsigned)port_);
    if (::getaddrinfo(ip_.c_str(), portstr, &hints, &res) != 0 || !res) throw rpc_error(0, "cannot resolve " + ip_);
    fd_ = ::socket(res->ai_family, res->ai_socktype, res->ai_protocol);
    const bool ok = fd_ >= 0 && ::connect(fd_, res->ai_addr, res->ai_addrlen) == 0;
    ::freeaddrinfo(res);
    if (!ok) { disconnect(); throw rpc_error(0, "cannot connect to " + ip_ + ":" + portstr); }
    int one = 1;
    ::setsockopt(fd_, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    timeval tv; tv.tv_sec = timeout_s_; tv.tv_usec = 0;
    ::setsockopt(fd_, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
  }
  // sends one request, waits for the response with the same msgid, returns its result or throws what its error says
  mp::Value call(const std::string& request) {
    ensure_connected();
    const uint32_t want = msgid_ - 1;
    if (!detail::send_all(fd_, request)) { disconnect(); throw rpc_error(0, "send failed"); }
    for (;;) {
      mp::Value msg;
      if (detail::recv_object(fd_, buf_, msg, (size_t)1 << 30) != 1) { disconnect(); throw rpc_error(0, "connection closed or timed out"); }
      if (!msg.is_array() || msg.a.size() != 4 || !msg.a[0].is_uint() || msg.a[0].u != RESPONSE || !msg.a[1].is_uint()) continue;
      if (msg.a[1].u != want) continue;                       // a late answer to an earlier, timed-out call
      const mp::Value& err = msg.a[2];
      if (err.type == mp::Value::NIL) return msg.a[3];
      if (err.is_uint()) throw rpc_error((int)err.u, err.u == NO_METHOD_ERROR ? "no such method" : err.u == ARGUMENT_ERROR ? "argument error" : "remote error");
      throw rpc_error(0, err.is_str() ? err.s : std::string("remote error"));
    }
  }
  std::string ip_;
  uint16_t port_;
  int timeout_s_;
  int fd_;
  uint32_t msgid_;
  std::string buf_;
};

}  // namespace vcrpc

#endif
