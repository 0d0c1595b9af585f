// hw5 <input> <output> — same CLI and files as the reference (hw5.cu:532-535, nbody.cc:91-94).
#include <cstdio>
#include <stdexcept>

#include "../../include/nbody_b200.h"

int main(int argc, char** argv) {
    if (argc != 3) {
        throw std::runtime_error("must supply 2 arguments");  // hw5.cu:533-535
    }
    int rc = nb_hw5_main(argv[1], argv[2], 0);
    if (rc != NB_OK) {
        fprintf(stderr, "hw5: %s: %s\n", nb_strerror(rc), nb_last_error_detail());
        return 1;
    }
    return 0;
}
