// worker_main.cpp -- `ompimg_worker`: the long-lived process that holds the CUDA context for the reference's
// exec-per-request services (see worker.cu).
//
//   ompimg_worker [--socket PATH] [--device N]        default socket: $OMPIMG_WORKER or /tmp/ompimg_b200.sock
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/ompimg_b200.h"

int main(int argc, char **argv) {
    const char *path = getenv("OMPIMG_WORKER");
    int device = -1;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--socket") && i + 1 < argc) path = argv[++i];
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else { fprintf(stderr, "usage: %s [--socket PATH] [--device N]\n", argv[0]); return 2; }
    }
    if (!path || !*path) path = "/tmp/ompimg_b200.sock";
    const int rc = ompimg_worker_serve(path, device);
    if (rc != OMPIMG_OK) { fprintf(stderr, "ompimg_worker: %s (%d)\n", ompimg_last_error(), rc); return 1; }
    return 0;
}
