// CPU-only check of halo2_b200_transcript.hpp: replays a list of transcript operations read from stdin and prints the
// challenges and the proof bytes, which tests/test_prover_host.py compares with the Python transcripts.
//   ops:  S <64 hex>   common_scalar (Montgomery limbs, little-endian bytes)      s <64 hex>  write_scalar
//         P <128 hex>  common_point (affine x || y, Fq Montgomery)               p <128 hex> write_point
//         C            squeeze_challenge -> prints "C <64 hex of the Montgomery limbs>"
#include <cstdio>
#include <iostream>
#include <string>

#include "halo2_b200_transcript.hpp"

using namespace halo2_b200;

static void unhex(const std::string& s, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)std::stoi(s.substr(2 * i, 2), nullptr, 16);
}

int main() {
    Blake2bWrite tr;
    std::string op, arg;
    while (std::cin >> op) {
        if (op == "C") {
            Fr c = tr.squeeze_challenge();
            const uint8_t* b = reinterpret_cast<const uint8_t*>(c.l);
            std::printf("C ");
            for (int i = 0; i < 32; i++) std::printf("%02x", b[i]);
            std::printf("\n");
            continue;
        }
        std::cin >> arg;
        if (op == "S" || op == "s") {
            Fr x;
            unhex(arg, reinterpret_cast<uint8_t*>(x.l), 32);
            if (op == "S") tr.common_scalar(x); else tr.write_scalar(x);
        } else {
            G1Affine p;
            unhex(arg, reinterpret_cast<uint8_t*>(&p), 64);
            try {
                if (op == "P") tr.common_point(p); else tr.write_point(p);
            } catch (const std::runtime_error& e) {
                std::printf("E %s\n", e.what());
            }
        }
    }
    std::printf("W ");
    for (uint8_t b : tr.finalize()) std::printf("%02x", b);
    std::printf("\n");
    return 0;
}
