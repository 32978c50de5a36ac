// TEST INFRASTRUCTURE ONLY (oracle).  Stand-in for the reference's Pilaf
// client (Pilaf/store-client.h:16-19,103-117,185): instead of an RDMA cuckoo
// DHT it is a process-wide in-memory byte-string KV store, so the reference's
// own PilafProxy (src/pilaf_proxy.h, compiled unmodified) can serialise keys
// and values and "store" them exactly as it would over Infiniband.  All
// Client objects in the process share one store, the way all reference
// processes share one Pilaf cluster.
#ifndef VC_ORACLE_SHIM_STORE_CLIENT_H
#define VC_ORACLE_SHIM_STORE_CLIENT_H
#include <stddef.h>
#include <stdint.h>

enum read_modes { READ_MODE_RDMA, READ_MODE_SERVER };
#define POST_GET_FOUND 0
#define POST_GET_MISSING 1

extern "C" void vc_shim_kv_clear(void);
extern "C" uint64_t vc_shim_kv_size(void);

class Client {
 public:
  int setup() { return 0; }
  int add_server(const char*, const char*) { return 0; }
  void set_read_mode(read_modes) {}
  int ready() { return 0; }
  void teardown() {}
  int put_with_size(const char* key, const char* value, size_t key_len, size_t val_len);
  int get_with_size(const char* key, char* value, size_t key_len, size_t& val_len);
};
#endif
