// gsc_decode -- command-line decoder with the reference's interface (dec:222-254):
//     gsc_decode <source.gsc> [dest.wav]
#include <cstdio>
#include <string>

#include "../include/gsc_host.h"

int main(int argc, char **argv) {
    if (argc < 2) { printf("Usage: %s <source GSC file> [dest WAV file]\n", argv[0]); return 0; }
    std::string dst = argc > 2 ? argv[2] : std::string(argv[1]) + ".wav";
    if (gsch_decode_file(argv[1], dst.c_str())) { fprintf(stderr, "error: %s\n", gsch_last_error()); return 1; }
    printf("Done: %s\n", dst.c_str());
    return 0;
}
