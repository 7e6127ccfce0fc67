// ASan/UBSan fuzz of the host bit reader on garbage / truncated streams.  Build + run (CPU only):
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -Iicspcodec_b200/host tools/host_fuzz_parse.cpp icspcodec_b200/host/bitstream.cpp -o /tmp/fuzz_parse -pthread && /tmp/fuzz_parse 1
#include "bitstream.h"
#include <cstdio>
#include <cstdlib>
#include <random>
int main(int argc, char** argv)
{
    std::mt19937 rng(argc > 1 ? atoi(argv[1]) : 1);
    int ok = 0, thrown = 0;
    for (int it = 0; it < 3000; it++) {
        const int w = 16 * (1 + rng() % 4), h = 16 * (1 + rng() % 3), ip = 1 + rng() % 5, n = 1 + rng() % 4;
        icsp_host::StreamParams p; p.width = w; p.height = h; p.qp_dc = 1 + rng() % 30; p.qp_ac = 1 + rng() % 30; p.intra_period = ip; p.nframes = n;
        std::vector<uint8_t> f = icsp_host::stream_header(p);
        const int body = rng() % 3000;
        const int mode = rng() % 4;
        for (int i = 0; i < body; i++) f.push_back(mode == 0 ? 0 : mode == 1 ? 0xff : mode == 2 ? (uint8_t)rng() : (uint8_t)((rng() % 4) ? 0 : rng()));
        if (rng() % 10 == 0) f.resize(rng() % 20);            // truncated header
        try { auto ps = icsp_host::parse_stream(f, n + rng() % 3); ok++; (void)ps; } catch (const std::exception&) { thrown++; }
    }
    printf("parsed %d, rejected %d\n", ok, thrown);
    return 0;
}
