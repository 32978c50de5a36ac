// image-search-server - the reference's query server (src/image_server_main.cc, src/image_search_server.{h,cc}) in
// front of the GPU-resident index: msgpack-rpc on the wire ("ping", "search_image_by_id"), answered in-process
// instead of by forking `ssh <worker> ... mpirun` per query and parsing its output (src/image_search_server.cc:58-100).
//
// Flags of the reference (src/image_server_main.cc:17-31): --port/-p (9191), --ip/-i (0.0.0.0), --nthreads/-n (10),
// --config_path/-c - there the list of worker hosts, here the device list of the GPU proxy (one CUDA ordinal per line;
// omitted = device 0).  Added, because the index now lives in this process: --index/-x <file written by build-tables -o>
// or --binary_file/-f <raw codes> with --binary_bits/-b, --ntables/-t, --image_total/-N.
// One search at a time reaches the GPU (the C ABI is single-threaded per index, include/verticut_gpu.h); the n
// threads overlap the network side.
#include <getopt.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>

#include <iostream>
#include <mutex>

#include "image_search_client.h"

static uint16_t port = 9191;                      // DEFAULT_SERVER_PORT, src/image_search_constants.h:16
static std::string ip = "0.0.0.0";
static const char* config_path = 0;
static int n_threads = 10;
static const char* index_in = 0;
static const char* binary_file = 0;
static int binary_bits = 128, n_tables = 4;       // N_BINARY_BITS, DEFAULT_N_TABLES
static long image_total = 100000000;              // DEFAULT_IMAGE_TOTAL
static vcrpc::image_search_server* server = 0;

static struct option long_options[] = {
    {"config_path", required_argument, 0, 'c'}, {"port", required_argument, 0, 'p'},        {"ip", required_argument, 0, 'i'},
    {"nthreads", required_argument, 0, 'n'},    {"index", required_argument, 0, 'x'},       {"binary_file", required_argument, 0, 'f'},
    {"binary_bits", required_argument, 0, 'b'}, {"ntables", required_argument, 0, 't'},     {"image_total", required_argument, 0, 'N'},
    {0, 0, 0, 0}};

static void usage() {
  fprintf(stderr,
          "image-search-server: msgpack-rpc query server in front of a GPU-resident index\n"
          "  -p, --port <n>          TCP port (default 9191; 0 = any free port, printed as \"port <n>\")\n"
          "  -i, --ip <addr>         address to bind (default 0.0.0.0)\n"
          "  -n, --nthreads <n>      connections served at the same time (default 10)\n"
          "  -c, --config_path <f>   device list of the GPU proxy, one CUDA ordinal per line (default: device 0)\n"
          "  -x, --index <f>         index written by build-tables -o, or instead:\n"
          "  -f, --binary_file <f>   raw code file, with -b/--binary_bits, -t/--ntables, -N/--image_total\n");
  exit(2);
}

struct gpu_service : vcrpc::service {
  explicit gpu_service(GpuTableProxy* proxy) : client(proxy) {}
  vcrpc::result_list search_image_by_id(uint32_t id, uint32_t knn, bool approximate) {
    // knn comes off the wire as a uint32: 0, > VC_MAX_K or >= 2^31 must not reach the library (one bad request would
    // otherwise take down the process that holds the only copy of the GPU index); the exception becomes the call's error
    if (knn == 0 || knn > (uint32_t)VC_MAX_K) throw std::runtime_error("knn must be in [1, " + std::to_string(VC_MAX_K) + "]");
    std::lock_guard<std::mutex> g(mu);
    vcrpc::result_list r = client.search_image_by_id(id, (int)knn, approximate);
    printf("finish query for %u\n", id);          // src/image_search_server.cc:82
    return r;
  }
  image_search_client client;
  std::mutex mu;
};

static void sig_handler(int) { _exit(1); }        // the reference closes and exits(1) on SIGINT as well (:69-77)

int main(int argc, char* argv[]) {
  signal(SIGINT, sig_handler);
  int opt, idx = 0;
  while ((opt = getopt_long(argc, argv, "c:p:i:n:x:f:b:t:N:", long_options, &idx)) != -1) {
    switch (opt) {
      case 'c': config_path = optarg; break;
      case 'p': port = (uint16_t)atoi(optarg); break;
      case 'i': ip = optarg; break;
      case 'n': n_threads = atoi(optarg); break;
      case 'x': index_in = optarg; break;
      case 'f': binary_file = optarg; break;
      case 'b': binary_bits = atoi(optarg); break;
      case 't': n_tables = atoi(optarg); break;
      case 'N': image_total = atol(optarg); break;
      default: usage();
    }
  }
  if (!index_in && !binary_file) usage();
  GpuTableProxy proxy(binary_bits, n_tables);
  if (proxy.init(config_path) != 0) { fprintf(stderr, "proxy init failed: %s\n", proxy.last_error()); return 1; }
  if (index_in) {
    if (proxy.load(index_in) != 0) { fprintf(stderr, "Can't read index %s\n", index_in); return 1; }
  } else {
    if (proxy.load_code_file(binary_file, (uint64_t)image_total) != 0) { fprintf(stderr, "Can't open file %s.\n", binary_file); return 1; }
    if (proxy.finalize() != 0) { fprintf(stderr, "build failed: %s\n", proxy.last_error()); return 1; }
  }
  gpu_service svc(&proxy);
  server = new vcrpc::image_search_server(&svc);
  const int bound = server->listen(ip, port);
  if (bound <= 0) { perror("listen"); return 1; }
  std::cout << "Server is running with " << n_threads << " threads..." << std::endl;   // src/image_server_main.cc:89
  std::cout << "port " << bound << std::endl;
  server->run(n_threads);
  return 0;
}
