// TEST INFRASTRUCTURE ONLY (oracle).  Declarations that let the reference's
// src/memcached_proxy.h parse and link without libmemcached.  The oracle never
// selects the memcached backend; every entry point aborts if reached.
#ifndef VC_ORACLE_SHIM_MEMCACHED_H
#define VC_ORACLE_SHIM_MEMCACHED_H
#include <stdlib.h>
#include <stdint.h>
#include <stdio.h>
#include <time.h>

struct memcached_st;
typedef int memcached_return_t;
typedef int memcached_behavior_t;
#define MEMCACHED_SUCCESS 0
#define MEMCACHED_DEFAULT_PORT 11211
#define MEMCACHED_BEHAVIOR_BINARY_PROTOCOL 18

static inline void vc_shim_no_memcached(void) {
  fprintf(stderr, "oracle shim: memcached backend is not available\n");
  abort();
}
static inline memcached_st* memcached_create(memcached_st*) { vc_shim_no_memcached(); return 0; }
static inline void memcached_free(memcached_st*) { vc_shim_no_memcached(); }
static inline memcached_return_t memcached_server_add(memcached_st*, const char*, int) { vc_shim_no_memcached(); return 1; }
static inline memcached_return_t memcached_behavior_set(memcached_st*, memcached_behavior_t, uint64_t) { vc_shim_no_memcached(); return 1; }
static inline memcached_return_t memcached_set(memcached_st*, const char*, size_t, const char*, size_t, time_t, uint32_t) { vc_shim_no_memcached(); return 1; }
static inline char* memcached_get(memcached_st*, const char*, size_t, size_t*, uint32_t*, memcached_return_t*) { vc_shim_no_memcached(); return 0; }
#endif
