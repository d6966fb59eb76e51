// Mutation fuzzer for agr_json_decode, built with -fsanitize=address,undefined by tests/test_json_decode.py.
// argv: seeds file (u32 length + bytes, repeated), iterations.  Exit code 0 = no sanitizer report (ASAN aborts otherwise).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include "agentainer_gpu.h"
int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb"); if (!f) return 2;
    const long iters = atol(argv[2]); std::vector<std::string> seeds;
    for (;;) { uint32_t n; if (fread(&n, 4, 1, f) != 1) break; std::string s(n, 0); if (fread(&s[0], 1, n, f) != n) return 2; seeds.push_back(s); }
    srand(7); long ok = 0, bad = 0, cap = 0;
    for (long it = 0; it < iters; ++it) {
        std::string b = seeds[rand() % seeds.size()];
        int k = rand() % 5;
        if (k == 0) for (int m = 0; m < 1 + rand() % 3; ++m) b[rand() % b.size()] = (char)(rand() % 256);
        else if (k == 1) b.resize(rand() % b.size());
        else if (k == 2) { size_t i = rand() % b.size(); std::string ins; for (int m = 0; m < 1 + rand() % 7; ++m) ins.push_back((char)(rand() % 256)); b.insert(i, ins); }
        else if (k == 3) { size_t i = rand() % b.size(); b.erase(i, 1 + rand() % 30); }
        else { size_t i = rand() % b.size(); b.insert(i, "\\ud83d\\ude00\\udc00\\u"); }
        // exact-size heap copies so that ASAN sees any over-read; small output buffers to exercise AGR_ECAP
        std::vector<uint8_t> in(b.begin(), b.end());
        const uint32_t rc_cap = (rand() % 4 == 0) ? 128 : 8192, rs_cap = (rand() % 4 == 0) ? 8 : 65536, er_cap = (rand() % 4 == 0) ? 2 : 4096;
        std::vector<uint8_t> rec(rc_cap), resp(rs_cap); std::vector<char> err(er_cap);
        agr_decoded d;
        int rc = agr_json_decode(in.data(), (uint32_t)in.size(), rec.data(), rc_cap, resp.data(), rs_cap, err.data(), er_cap, &d);
        if (rc == 0) ok++; else if (rc == AGR_ECAP) cap++; else bad++;
    }
    printf("ok %ld ecap %ld einval %ld\n", ok, cap, bad);
    return 0;
}
