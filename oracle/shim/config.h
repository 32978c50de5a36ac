// TEST INFRASTRUCTURE ONLY (oracle).  Stand-in for the reference's
// Pilaf/config.h:6-63 (ConfigReader: "host port" lines) so src/pilaf_proxy.h
// parses.  Same public surface; it reads the same file format.
#ifndef VC_ORACLE_SHIM_CONFIG_H
#define VC_ORACLE_SHIM_CONFIG_H
#include <stdio.h>
#include <string>
#include <vector>

struct server_info {
  std::string* host;
  std::string* port;
};

class ConfigReader {
  std::vector<server_info*> servers_;
  size_t at_;
 public:
  explicit ConfigReader(std::string fname) : at_(0) {
    FILE* fh = fopen(fname.c_str(), "r");
    if (!fh) return;
    char host[1024], port[1024];
    while (fscanf(fh, "%1023s %1023s", host, port) == 2) {
      server_info* s = new server_info;
      s->host = new std::string(host);
      s->port = new std::string(port);
      servers_.push_back(s);
    }
    fclose(fh);
  }
  int get_count() { return (int)servers_.size(); }
  server_info* get_next() { return servers_[at_++]; }
  void get_reset() { at_ = 0; }
  bool get_end() { return at_ >= servers_.size(); }
};
#endif
