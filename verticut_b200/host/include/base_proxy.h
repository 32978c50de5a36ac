// Key-value proxy interface of the MIH path, as the reference declares it (src/base_proxy.h:10-29):
// the search core and the table builder talk to storage only through get/put/contain/init/close on a
// BaseProxy<Message, Message>, and select the backend at run time.  GpuTableProxy (gpu_table_proxy.h)
// is the backend this project adds; this header restates the interface so that code written against
// the reference's proxies compiles against it unchanged.
#ifndef VERTICUT_B200_BASE_PROXY_H
#define VERTICUT_B200_BASE_PROXY_H

// return codes of get / put (src/base_proxy.h:10-13)
#define PROXY_FOUND 0
#define PROXY_NOT_FOUND 1
#define PROXY_PUT_DONE 0
#define PROXY_PUT_FAIL 1

template <class K, class V>
class BaseProxy {
 public:
  virtual ~BaseProxy() {}
  // PROXY_FOUND and `value` filled, or PROXY_NOT_FOUND
  virtual int get(const K& key, V& value) = 0;
  // PROXY_PUT_DONE / PROXY_PUT_FAIL
  virtual int put(const K& key, const V& value) = 0;
  virtual int contain(const K& key) = 0;
  // `filename` lists the servers of the backend, one per line; 0 on success, -1 on error
  virtual int init(const char* filename) = 0;
  virtual void close() = 0;
};

#endif
