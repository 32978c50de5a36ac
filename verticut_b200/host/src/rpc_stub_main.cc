// rpc-stub - the msgpack-rpc front end (image_search_rpc.h) in front of a deterministic stand-in service: what the wire-format
// tests run on machines without a GPU (tests/test_rpc_wire.py talks to it with the third-party `msgpack` package).
//   rpc-stub serve [port]     listens on 127.0.0.1 (port 0 = any), prints "port <n>", serves until stdin closes
//   rpc-stub selftest         server + remote_client in one process; prints "ok" or the first mismatch
// The stand-in answers search_image_by_id(id, knn, approximate) with knn pairs (id + i, knn - 1 - i [+ 1000 if approximate]) -
// descending distance, the order the real service returns - and throws "Can't find match" for id 0xFFFFFFFF, the
// message the real one throws for an unknown id (image_search_client.h).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "image_search_rpc.h"

struct stub_service : vcrpc::service {
  vcrpc::result_list search_image_by_id(uint32_t id, uint32_t knn, bool approximate) {
    if (id == 0xFFFFFFFFu) throw std::runtime_error("Can't find match");
    vcrpc::result_list r;
    for (uint32_t i = 0; i < knn; ++i) r.push_back(std::make_pair(id + i, knn - 1 - i + (approximate ? 1000u : 0u)));
    return r;
  }
};

#define CHECK(cond) do { if (!(cond)) { printf("FAILED: %s (line %d)\n", #cond, __LINE__); return 1; } } while (0)

static int selftest() {
  // wire format: bytes worked out by hand from the MessagePack specification
  CHECK(vcrpc::request_ping(1, "hi") == std::string("\x94\x00\x01\xa4ping\x91\xa2hi", 12));
  CHECK(vcrpc::request_search(7, 300, 70000, true) ==
        std::string("\x94\x00\x07\xb2search_image_by_id\x93\xcd\x01\x2c\xce\x00\x01\x11\x70\xc3", 32));
  {
    mp::Packer pk;
    pk.pack_int(-1); pk.pack_int(-33); pk.pack_int(-129); pk.pack_uint(255); pk.pack_uint(65536); pk.pack_uint(1ull << 32);
    pk.pack_str(std::string(32, 'x'));
    const std::string want = std::string("\xff\xd0\xdf\xd1\xff\x7f\xcc\xff\xce\x00\x01\x00\x00\xcf\x00\x00\x00\x01\x00\x00\x00\x00\xda\x00\x20", 25) + std::string(32, 'x');
    CHECK(pk.bytes() == want);
    size_t pos = 0, used = 0;
    mp::Value v;
    const int64_t ints[3] = {-1, -33, -129};
    for (int i = 0; i < 3; ++i) { CHECK(mp::parse(want.data() + pos, want.size() - pos, used, v) == mp::PARSE_OK); CHECK(v.type == mp::Value::INT && v.i == ints[i]); pos += used; }
    const uint64_t uints[3] = {255, 65536, 1ull << 32};
    for (int i = 0; i < 3; ++i) { CHECK(mp::parse(want.data() + pos, want.size() - pos, used, v) == mp::PARSE_OK); CHECK(v.type == mp::Value::UINT && v.u == uints[i]); pos += used; }
    CHECK(mp::parse(want.data() + pos, want.size() - pos, used, v) == mp::PARSE_OK && v.is_str() && v.s == std::string(32, 'x'));
    CHECK(pos + used == want.size());
    // every proper prefix of a message is "need more", never an error or a short object
    const std::string req = vcrpc::request_search(7, 300, 70000, true);
    for (size_t n = 0; n < req.size(); ++n) CHECK(mp::parse(req.data(), n, used, v) == mp::PARSE_NEED_MORE);
    CHECK(mp::parse(req.data(), req.size(), used, v) == mp::PARSE_OK && used == req.size());
    CHECK(mp::parse("\xc1", 1, used, v) == mp::PARSE_ERROR);
  }
  stub_service svc;
  // dispatch without sockets
  {
    mp::Value msg, resp;
    size_t used = 0;
    const std::string req = vcrpc::request_ping(5, "abc");
    CHECK(mp::parse(req.data(), req.size(), used, msg) == mp::PARSE_OK);
    CHECK(vcrpc::handle_message(svc, msg) == std::string("\x94\x01\x05\xc0\xa3" "abc", 8));
    const std::string bad("\x94\x00\x09\xa3" "foo" "\x90", 8);                        // unknown method
    CHECK(mp::parse(bad.data(), bad.size(), used, msg) == mp::PARSE_OK);
    CHECK(vcrpc::handle_message(svc, msg) == std::string("\x94\x01\x09\x01\xc0", 5));       // NO_METHOD_ERROR
    const std::string arg("\x94\x00\x0a\xa4ping\x91\x07", 10);                          // ping(7): not a string
    CHECK(mp::parse(arg.data(), arg.size(), used, msg) == mp::PARSE_OK);
    CHECK(vcrpc::handle_message(svc, msg) == std::string("\x94\x01\x0a\x02\xc0", 5));       // ARGUMENT_ERROR
    const std::string note("\x93\x02\xa4ping\x91\xa1x", 10);                            // notification: no answer
    CHECK(mp::parse(note.data(), note.size(), used, msg) == mp::PARSE_OK);
    CHECK(vcrpc::handle_message(svc, msg).empty());
  }
  // through TCP
  vcrpc::image_search_server server(&svc);
  const int port = server.listen("127.0.0.1", 0);
  CHECK(port > 0);
  server.start(3);
  {
    vcrpc::remote_client c("127.0.0.1", (uint16_t)port, 10), c2("127.0.0.1", (uint16_t)port, 10);
    CHECK(c.ping("hello") == "hello");
    CHECK(c.ping(std::string(100000, 'y')) == std::string(100000, 'y'));               // raw32 both ways
    vcrpc::result_list r = c.search_image_by_id(40, 5, false);
    CHECK(r.size() == 5 && r.front() == std::make_pair(40u, 4u) && r.back() == std::make_pair(44u, 0u));
    r = c2.search_image_by_id(4000000000u, 3, true);
    CHECK(r.size() == 3 && r.front() == std::make_pair(4000000000u, 1002u));
    r = c.search_image_by_id(1, 20000, false);                                         // array32 of pairs
    CHECK(r.size() == 20000 && r.back() == std::make_pair(20000u, 0u));
    bool threw = false;
    try { c.search_image_by_id(0xFFFFFFFFu, 3, false); } catch (const vcrpc::rpc_error& e) { threw = e.code == 0 && std::string(e.what()) == "Can't find match"; }
    CHECK(threw);
    CHECK(c.ping("still alive") == "still alive");                                     // the connection survives a remote error
  }
  server.stop();
  printf("ok\n");
  return 0;
}

int main(int argc, char* argv[]) {
  if (argc >= 2 && !strcmp(argv[1], "selftest")) return selftest();
  if (argc >= 2 && !strcmp(argv[1], "serve")) {
    stub_service svc;
    vcrpc::image_search_server server(&svc);
    const int port = server.listen("127.0.0.1", argc >= 3 ? (uint16_t)atoi(argv[2]) : 0);
    if (port <= 0) { perror("listen"); return 1; }
    printf("port %d\n", port);
    fflush(stdout);
    server.start(4);
    char buf[64];
    while (fgets(buf, sizeof buf, stdin)) {}                                           // serve until the parent closes stdin
    server.stop();
    return 0;
  }
  fprintf(stderr, "usage: rpc-stub serve [port] | rpc-stub selftest\n");
  return 2;
}
