// Command-line flags of the reference's tools (src/args_config.{h,cc}, defaults src/image_search_constants.h:6-18):
//   --server/-s  --config_path/-c  --binary_bits/-b  --read_mode/-r  --ntables/-n  --binary_file/-f  -i  -k  --help/-h
// plus the values this project adds: server "gpu" (default) and, for the search tools, -q <query file> / -a.
#ifndef VERTICUT_B200_ARGS_CONFIG_H
#define VERTICUT_B200_ARGS_CONFIG_H

extern const char* config_path;   // device-list file for --server gpu (one CUDA ordinal per line); NULL = device 0
extern const char* server;        // "gpu"; the reference's "pilaf" / "memcached" / "redis" are refused
extern const char* binary_file;   // raw code file, record = binary_bits/8 bytes, id = ordinal
extern const char* query_file;    // raw query file, same record format
extern int binary_bits;           // 128
extern int n_tables;              // 4
extern int read_mode;             // accepted and ignored (Pilaf only)
extern int image_total;           // 100000000; upper bound on the codes read from binary_file
extern int knn;                   // 10
extern int approximate;           // -a
extern int max_queries;           // 200 (src/distributed_image_search.cc:83)
extern const char* index_out;     // -o: build-tables writes the built index here
extern const char* index_in;      // -x: search tools read a built index instead of building from binary_file
extern int query_image_id;        // -I, byid mode (argv[9] of src/distributed_image_search.cc:153)

void configure(int argc, char* argv[]);
void usage();

#endif
