#include "args_config.h"

#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

const char* config_path = 0;
const char* server = "gpu";
const char* binary_file = "lsh.code";
const char* query_file = 0;
int binary_bits = 128;
int n_tables = 4;
int read_mode = 0;
int image_total = 100000000;
int knn = 10;
int approximate = 0;
int max_queries = 200;
int query_image_id = -1;
const char* index_out = 0;
const char* index_in = 0;

static struct option long_options[] = {
    {"server", required_argument, 0, 's'},      {"config_path", required_argument, 0, 'c'},
    {"binary_bits", required_argument, 0, 'b'}, {"read_mode", required_argument, 0, 'r'},
    {"ntables", required_argument, 0, 'n'},     {"binary_file", required_argument, 0, 'f'},
    {"query_file", required_argument, 0, 'q'},  {"approximate", no_argument, 0, 'a'},
    {"help", no_argument, 0, 'h'},              {0, 0, 0, 0}};

void usage() {
  printf("Usage : \n");
  printf("--server -s : key-value backend holding the tables. [gpu]\n");
  printf("--config_path -c : file listing the CUDA device ordinal(s), one per line (default: device 0).\n");
  printf("--binary_bits -b : How many bits of each binary code.\n");
  printf("--ntables -n : How many sub-tables we use.\n");
  printf("--binary_file -f : The path of the binary file. \n");
  printf("--query_file -q : raw query codes, same record format as the binary file.\n");
  printf("-i : The number of images the server has (upper bound on the codes read).\n");
  printf("-k : Find k nearest neighbors.\n");
  printf("-a : approximate search (factor 20).\n");
  printf("-I : image id to query by (byid mode).\n");
  printf("-o : build-tables: write the built index to this file.  -x : search tools: read the index from this file instead of building it from -f.\n");
  printf("-r : The read mode (accepted for compatibility, Pilaf only).\n");
  printf("--help -h : help information.\n");
  exit(-1);
}

void configure(int argc, char* argv[]) {
  int opt, opt_index = 0;
  optind = 1;
  while ((opt = getopt_long(argc, argv, "c:b:r:n:s:i:k:f:q:I:o:x:ah", long_options, &opt_index)) != -1) {
    switch (opt) {
      case 'n': n_tables = atoi(optarg); break;
      case 's': server = optarg; break;
      case 'b': binary_bits = atoi(optarg); break;
      case 'c': config_path = optarg; break;
      case 'r': read_mode = atoi(optarg); break;
      case 'i': image_total = atoi(optarg); break;
      case 'k': knn = atoi(optarg); break;
      case 'f': binary_file = optarg; break;
      case 'q': query_file = optarg; break;
      case 'a': approximate = 1; break;
      case 'I': query_image_id = atoi(optarg); break;
      case 'o': index_out = optarg; break;
      case 'x': index_in = optarg; break;
      default: usage();
    }
  }
  if (strcmp(server, "gpu") != 0) {
    fprintf(stderr, "Unknown server type '%s': this build keeps the tables in GPU memory (--server gpu).\n", server);
    usage();
  }
}
