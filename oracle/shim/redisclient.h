// TEST INFRASTRUCTURE ONLY (oracle).  Declarations that let the reference's
// src/redis_proxy.h parse and link without boost / a redis server.  The oracle
// never selects the redis backend; every entry point aborts if reached.
#ifndef VC_ORACLE_SHIM_REDISCLIENT_H
#define VC_ORACLE_SHIM_REDISCLIENT_H
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
namespace redis {
struct connection_data {
  std::string host;
  int port;
  int dbindex;
  connection_data() : port(0), dbindex(0) {}
};
class client {
 public:
  template <class It> client(It, It) { fprintf(stderr, "oracle shim: redis backend is not available\n"); abort(); }
  void set(const std::string&, const std::string&) { abort(); }
  std::string get(const std::string&) { abort(); return std::string(); }
};
}  // namespace redis
#endif
