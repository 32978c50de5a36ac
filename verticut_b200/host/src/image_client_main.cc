// image-search-client - command-line front of the reference's image_search_client (src/image_search_client.h:12-26)
// over msgpack-rpc:   image-search-client <ip> <port> ping <text>
//                     image-search-client <ip> <port> byid <image id> <knn> [-a]
// Result lines "id : dist", the format of the reference's search tools (src/image_search_server.cc:94 parses it).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <iostream>

#include "image_search_rpc.h"

int main(int argc, char* argv[]) {
  if (argc < 5) { fprintf(stderr, "usage: image-search-client <ip> <port> ping <text> | byid <image id> <knn> [-a]\n"); return 2; }
  vcrpc::remote_client client(argv[1], (uint16_t)atoi(argv[2]));
  try {
    if (!strcmp(argv[3], "ping")) {
      std::cout << client.ping(argv[4]) << std::endl;
    } else if (!strcmp(argv[3], "byid") && argc >= 6) {
      const bool approximate = argc >= 7 && !strcmp(argv[6], "-a");
      const vcrpc::result_list r = client.search_image_by_id((uint32_t)strtoul(argv[4], 0, 10), atoi(argv[5]), approximate);
      for (vcrpc::result_list::const_iterator it = r.begin(); it != r.end(); ++it) std::cout << it->first << " : " << it->second << std::endl;
    } else {
      fprintf(stderr, "unknown command %s\n", argv[3]);
      return 2;
    }
  } catch (const vcrpc::rpc_error& e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
