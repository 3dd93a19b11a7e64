// AddressSanitizer harness for the proof (de)serialiser (host/serialize.cu): a serialised proof (hex on stdin) is mutated
// N times (byte flips, truncations, insertions, header edits); every mutant goes through lsp_proof_deserialize with
// exact-size heap buffers, and whatever is accepted must serialise back to the same bytes.
//   nvcc -O1 -g -std=c++17 -Xcompiler -fsanitize=address,-fno-omit-frame-pointer -o ser_asan linea-stark-prover_b200/host/serialize.cu \
//        tools/serialize_asan_harness.cpp -lcudart      (lsp_proof_words is restated below: prover.cu is device code)
//   python -c "import json;print(json.load(open('tests/golden/golden_round2_v1.json'))['serialized_hex'])" | ./ser_asan 200000
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "../include/lsp_b200.h"

extern "C" size_t lsp_proof_words(uint32_t log_n, uint32_t width, uint32_t log_q, const lsp_fri_config* fri) {  // as host/prover.cu
    if (!fri || fri->log_final_poly_len >= log_n) return 0;
    size_t log_l = size_t(log_n) + fri->log_blowup, q = size_t(1) << log_q, rounds = log_n - fri->log_final_poly_len;
    size_t f = size_t(1) << (fri->log_blowup + fri->log_final_poly_len);
    size_t per_query = 1 + (width + log_l) + (q + log_l);
    for (size_t r = 0; r < rounds; r++) per_query += 1 + (log_l - 1 - r);
    return (2 + 2 * size_t(width) + q + rounds + f + 1 + size_t(fri->num_queries) * per_query) * 4;
}

int main(int argc, char** argv) {
    const long n_mut = argc > 1 ? atol(argv[1]) : 100000;
    std::string hex;
    char buf[1 << 16];
    while (size_t k = fread(buf, 1, sizeof buf, stdin)) hex.append(buf, k);
    std::vector<uint8_t> good;
    for (size_t i = 0; i + 1 < hex.size() && isxdigit(hex[i]); i += 2) good.push_back(uint8_t(strtol(hex.substr(i, 2).c_str(), nullptr, 16)));
    std::mt19937_64 rng(2026);
    long accepted = 0, rejected = 0;
    for (long it = 0; it <= n_mut; it++) {
        std::vector<uint8_t> m = good;
        if (it) switch (rng() % 5) {
            case 0: for (int k = 0; k < 1 + int(rng() % 3); k++) m[rng() % m.size()] = uint8_t(rng()); break;
            case 1: m.resize(rng() % m.size()); break;
            case 2: m.insert(m.begin() + rng() % m.size(), uint8_t(rng())); break;
            case 3: m[4 + rng() % 28] = uint8_t(rng()); break;                                         // header fields
            default: { size_t i = 32 + rng() % (m.size() - 40); uint64_t v = rng() % 3 ? rng() % 64 : rng(); memcpy(&m[i], &v, 8); }   // a length
        }
        uint8_t* blob = (uint8_t*)malloc(m.size() ? m.size() : 1);
        memcpy(blob, m.data(), m.size());
        uint32_t log_n = 0, width = 0, log_q = 0;
        lsp_fri_config fri;
        size_t words = 0;
        if (lsp_proof_deserialize(blob, m.size(), &log_n, &width, &log_q, &fri, nullptr, 0, &words) == 0 && words < (size_t(1) << 26)) {
            uint64_t* flat = (uint64_t*)malloc(words * 8);
            if (lsp_proof_deserialize(blob, m.size(), &log_n, &width, &log_q, &fri, flat, words, &words) == 0) {
                accepted++;
                std::vector<uint8_t> back(lsp_proof_serialized_bytes(log_n, width, log_q, &fri));
                size_t n = 0;
                if (lsp_proof_serialize(flat, words, log_n, width, log_q, &fri, back.data(), back.size(), &n) != 0 || n != m.size() ||
                    memcmp(back.data(), blob, n) != 0) {
                    printf("ROUND-TRIP MISMATCH at mutant %ld\n", it);
                    return 1;
                }
            } else rejected++;
            free(flat);
        } else rejected++;
        free(blob);
    }
    printf("mutants %ld: accepted %ld (each re-serialised to the same bytes), rejected %ld\n", n_mut, accepted, rejected);
    return 0;
}
