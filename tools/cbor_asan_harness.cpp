// AddressSanitizer harness for the CBOR reader (host/cbor.cu): every entry point, serial and with the parallel pre-pass
// forced on, over files given on the command line (each copied into an exact-size heap block so that a read past the end
// is caught).  Inputs: tests/test_cbor_fuzz.py:_mutations writes them (+ deep-nesting files); last run (round 2, incl. the
// padded `_decode` / `_read_rows` entry points and 100 000-deep nested values): see profiles/r3q_cbor_asan.log.
//   nvcc -O1 -g -std=c++17 -Xcompiler -fsanitize=address,-fno-omit-frame-pointer -o cbor_asan \
//        linea-stark-prover_b200/host/cbor.cu tools/cbor_asan_harness.cpp -lcudart
//   ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 ./cbor_asan in/*.bin
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>
#include "../include/lsp_b200.h"
int main(int argc, char** argv) {
    size_t ok = 0, bad = 0;
    for (int i = 1; i < argc; i++) {
        std::ifstream f(argv[i], std::ios::binary);
        std::vector<uint8_t> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        // exact-size heap copy so that ASan sees any read past the end
        uint8_t* blob = (uint8_t*)malloc(raw.size() ? raw.size() : 1);
        memcpy(blob, raw.data(), raw.size());
        for (int mode = 0; mode < 2; mode++) {
            setenv("LSP_CBOR_THREADS", mode ? "4" : "1", 1);
            if (mode) setenv("LSP_CBOR_PRESCAN_MIN", "0", 1); else unsetenv("LSP_CBOR_PRESCAN_MIN");
            size_t rows = 0; uint32_t a = 0, t = 0, b = 0; char name[64]; uint8_t* out = nullptr;
            int rc = lsp_cbor_permutation_read(blob, raw.size(), &rows, &a, name, sizeof name, &out);
            if (rc == 0) { ok++; lsp_host_free(out); } else bad++;
            out = nullptr;
            rc = lsp_cbor_lookup_read(blob, raw.size(), &rows, &a, &t, &b, name, sizeof name, &out);
            if (rc == 0) { ok++; lsp_host_free(out); } else bad++;
            size_t r2 = 0; uint32_t c2 = 0;
            if (lsp_cbor_permutation_shape(blob, raw.size(), &r2, &c2, name, sizeof name) == 0 && r2 && c2) {
                std::vector<uint8_t> o(r2 * 2 * c2 * 32);
                lsp_cbor_permutation_decode(blob, raw.size(), o.data(), r2, c2);
                // push_traces' resize: a taller common height (exact-size buffer: a write past the padded rows is caught)
                std::vector<uint8_t> taller((r2 + 5) * 2 * c2 * 32);
                lsp_cbor_permutation_decode(blob, raw.size(), taller.data(), r2 + 5, c2);
            }
            out = nullptr;
            if (lsp_cbor_permutation_read_rows(blob, raw.size(), 37, &rows, &a, name, sizeof name, &out) == 0) lsp_host_free(out);
            out = nullptr;
            if (lsp_cbor_lookup_read_rows(blob, raw.size(), 37, &rows, &a, &t, &b, name, sizeof name, &out) == 0) lsp_host_free(out);
        }
        free(blob);
    }
    printf("files %d: accepted %zu rejected %zu\n", argc - 1, ok, bad);
    return 0;
}
