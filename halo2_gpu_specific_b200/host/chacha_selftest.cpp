// Host build of csrc/chacha.cuh (the same source the device compiles): prints key stream blocks as hex, one per line.
//   chacha_selftest <64 hex digits key> <first counter> <count> [n0 n1 n2]
// tests/test_prover_host.py compares the output with RFC 8439 2.3.2 and with an independent ChaCha20.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../csrc/chacha.cuh"

int main(int argc, char** argv) {
    if (argc < 4 || strlen(argv[1]) != 64) {
        fprintf(stderr, "usage: %s <key: 64 hex digits> <first counter> <count> [n0 n1 n2]\n", argv[0]);
        return 2;
    }
    uint32_t key[8];
    for (int w = 0; w < 8; w++) {
        uint32_t v = 0;
        for (int b = 0; b < 4; b++) {
            unsigned byte = 0;
            sscanf(argv[1] + 8 * w + 2 * b, "%2x", &byte);
            v |= (uint32_t)byte << (8 * b);          // little-endian words, as RFC 8439 lays the key out
        }
        key[w] = v;
    }
    const uint32_t first = (uint32_t)strtoul(argv[2], nullptr, 0), count = (uint32_t)strtoul(argv[3], nullptr, 0);
    uint32_t n[3] = {0, 0, 0};
    for (int i = 0; i < 3 && 4 + i < argc; i++) n[i] = (uint32_t)strtoul(argv[4 + i], nullptr, 0);
    for (uint32_t c = 0; c < count; c++) {
        uint32_t out[16];
        b2::chacha20_block(key, first + c, n[0], n[1], n[2], out);
        for (int w = 0; w < 16; w++)
            for (int b = 0; b < 4; b++) printf("%02x", (out[w] >> (8 * b)) & 0xffu);
        printf("\n");
    }
    return 0;
}
