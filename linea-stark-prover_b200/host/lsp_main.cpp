// lsp_prove -- the reference's `main` (bin/src/main.rs:19-97) on top of the C ABI, in plain C++.
//
// What bin/src/main.rs does, step for step, with every field operation on the device:
//   challenges alpha, delta           rng.sample(Standard)                   :29-33   -> seeded draw below
//   RawTrace::new([alpha, delta])                                           :35
//   RawLookupTrace::read_file(..)     CBOR input files                       :37-41   -> lsp_cbor_*_shape/_decode
//   raw_trace.push_traces(perm, lookup)   lookups first, configs shifted     :43      -> lsp_lookup_trace_be,
//                                                                                        lsp_permutation_trace_be, lsp_mat_hconcat
//   Perm::new_from_rng(8, 22, rng), Hash, Compress, Mmcs, FriConfig, Pcs     :49-68   -> lsp_set_poseidon2, lsp_fri_config
//   LineaAIR::new(cfgs); prove(..)                                           :76-86   -> lsp_prove_air_dev
//   verify(..)                                                               :88-96   -> lsp_verify_air (the proof is also
//                                                                                        written to --out for any other verifier)
// This file includes only include/lsp_b200.h and the C++ standard library: it is the binding a host
// written in a compiled language would use, and it cannot fall back to anything -- no device, no proof.
//
//   lsp_prove [--lookup f.cbor]... [--permutation f.cbor]... [--seed S] [--log-blowup 3] [--queries 33]
//             [--pow-bits 0] [--sbox-d 5] [--device 0] [--out proof.bin] [--repeat 1]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/lsp_b200.h"

namespace {

// r, little-endian u64 limbs: any 4 x u64 below it is the Montgomery representative of a uniform field element
const uint64_t R_MOD[4] = {0x0a11800000000001ull, 0x59aa76fed0000001ull, 0x60b44d1e5c37b001ull, 0x12ab655e9a2ca556ull};
const uint64_t ONE_MONT[4] = {0x7d1c7ffffffffff3ull, 0x7257f50f6ffffff2ull, 0x16d81575512c0feeull, 0x0d4bda322bbb9a9dull};
const uint64_t TWO_MONT[4] = {0xf0277fffffffffe5ull, 0x8b0573200fffffe3ull, 0xccfbddcc46206fdbull, 0x07ec4f05bd4a8fe3ull};

struct SplitMix {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    void fr(uint64_t out[4]) {  // rejection sampling below r
        for (;;) {
            for (int i = 0; i < 4; i++) out[i] = next();
            out[3] &= (1ull << 61) - 1;
            for (int i = 3; i >= 0; i--) {
                if (out[i] < R_MOD[i]) return;
                if (out[i] > R_MOD[i]) break;
            }
        }
    }
};

// The input file, mapped read-only: the CBOR reader parses it in place (no copy of the ~50 bytes per element).
struct MappedFile {
    const uint8_t* p = nullptr;
    size_t n = 0;
    explicit MappedFile(const std::string& path) {
        int fd = open(path.c_str(), O_RDONLY);
        struct stat st;
        if (fd < 0 || fstat(fd, &st) != 0) {
            fprintf(stderr, "cannot open %s\n", path.c_str());
            exit(2);
        }
        n = size_t(st.st_size);
        if (n) {
            void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
            if (m == MAP_FAILED) {
                fprintf(stderr, "cannot map %s\n", path.c_str());
                exit(2);
            }
            p = static_cast<const uint8_t*>(m);
        }
        close(fd);
    }
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    ~MappedFile() {
        if (p) munmap(const_cast<uint8_t*>(p), n);
    }
    const uint8_t* data() const { return p; }
    size_t size() const { return n; }
};

// Owner of a buffer the library allocated (lsp_cbor_*_read).
struct HostBytes {
    uint8_t* p;
    explicit HostBytes(uint8_t* q) : p(q) {}
    HostBytes(const HostBytes&) = delete;
    HostBytes& operator=(const HostBytes&) = delete;
    ~HostBytes() { lsp_host_free(p); }
    uint8_t* data() { return p; }
};

void die(lsp_ctx* ctx, const char* what, int rc) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, ctx ? lsp_last_error(ctx) : "");
    exit(1);
}
#define CHECK(ctx, call)                     \
    do {                                     \
        int rc__ = (call);                   \
        if (rc__ != 0) die(ctx, #call, rc__); \
    } while (0)

struct LookupIds {  // storage behind one lsp_lookup_air_cfg
    std::vector<uint32_t> a, b, bf, bi, occ;
};

}  // namespace

int main(int argc, char** argv) {
    std::vector<std::string> lookups, perms;
    std::string out_path;
    uint64_t seed = 0xB200;
    int device = 0, sbox_d = 5, repeat = 1;
    lsp_fri_config fri = {3, 0, 33, 0};  // bin/src/main.rs:58-64
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](const char* name) -> const char* {
            if (i + 1 >= argc) {
                fprintf(stderr, "%s needs a value\n", name);
                exit(2);
            }
            return argv[++i];
        };
        if (a == "--lookup") lookups.push_back(need("--lookup"));
        else if (a == "--permutation") perms.push_back(need("--permutation"));
        else if (a == "--seed") seed = strtoull(need("--seed"), nullptr, 0);
        else if (a == "--log-blowup") fri.log_blowup = uint32_t(atoi(need("--log-blowup")));
        else if (a == "--queries") fri.num_queries = uint32_t(atoi(need("--queries")));
        else if (a == "--pow-bits") fri.proof_of_work_bits = uint32_t(atoi(need("--pow-bits")));
        else if (a == "--sbox-d") sbox_d = atoi(need("--sbox-d"));
        else if (a == "--device") device = atoi(need("--device"));
        else if (a == "--out") out_path = need("--out");
        else if (a == "--repeat") repeat = atoi(need("--repeat"));
        else {
            fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        }
    }
    if (lookups.empty() && perms.empty()) {
        fprintf(stderr, "usage: lsp_prove [--lookup f.cbor]... [--permutation f.cbor]... [--seed S] [--out proof.bin]\n");
        return 2;
    }

    lsp_ctx* ctx = nullptr;
    int rc = lsp_ctx_create(device, &ctx);
    if (rc != 0) {
        fprintf(stderr, "lsp_ctx_create(%d) failed (%d): a CUDA device is required, there is no CPU path\n", device, rc);
        return 1;
    }
    SplitMix rng{seed};
    uint64_t publics[2][4];
    rng.fr(publics[0]);  // alpha  (main.rs:30)
    rng.fr(publics[1]);  // delta  (main.rs:31)
    printf("Challenge delta: 0x%016llx%016llx%016llx%016llx (Montgomery limbs)\n", (unsigned long long)publics[1][3],
           (unsigned long long)publics[1][2], (unsigned long long)publics[1][1], (unsigned long long)publics[1][0]);
    printf("Challenge alpha: 0x%016llx%016llx%016llx%016llx (Montgomery limbs)\n", (unsigned long long)publics[0][3],
           (unsigned long long)publics[0][2], (unsigned long long)publics[0][1], (unsigned long long)publics[0][0]);

    // ---- Perm::new_from_rng(8, 22, &mut rng) (main.rs:49): 4x3 initial, 4x3 terminal, 22 internal constants
    std::vector<uint64_t> consts((8 * 3 + 22) * 4);
    for (size_t i = 0; i < consts.size() / 4; i++) rng.fr(&consts[4 * i]);
    uint64_t diag[3][4];
    memcpy(diag[0], ONE_MONT, 32);
    memcpy(diag[1], ONE_MONT, 32);
    memcpy(diag[2], TWO_MONT, 32);
    CHECK(ctx, lsp_set_poseidon2(ctx, 3, sbox_d, 8, 22, consts.data(), &diag[0][0]));

    // One pass of main's body.  --repeat N runs it N times in one process: the first pass pays for the CUDA context,
    // the memory pool and the twiddle / selector tables; later passes are what a long-running prover sees.
    auto run_once = [&]() -> int {
        // ---- push_traces: lookups first, then permutations; configs shifted by the running width (trace/src/lib.rs:62-92)
        printf("Generating trace...\n");
        using clk = std::chrono::steady_clock;
        auto ms_since = [](clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); };
        double t_read = 0, t_cbor = 0, t_witness = 0;
        size_t cbor_bytes = 0;
        const clk::time_point t_gen0 = clk::now();
        std::vector<lsp_mat*> parts;
        std::vector<lsp_lookup_air_cfg> lcfg;
        std::vector<lsp_perm_air_cfg> pcfg;
        std::vector<LookupIds> lids(lookups.size());
        std::vector<std::vector<uint32_t>> pa(perms.size()), pb(perms.size());
        uint32_t col = 0;
        size_t height = 0;
        for (size_t k = 0; k < lookups.size(); k++) {
            clk::time_point t0 = clk::now();
            MappedFile blob(lookups[k]);
            t_read += ms_since(t0);
            cbor_bytes += blob.size();
            size_t rows = 0;
            uint32_t na = 0, nt = 0, nb = 0;
            char name[128];
            t0 = clk::now();
            uint8_t* be_raw = nullptr;
            CHECK(ctx, lsp_cbor_lookup_read(blob.data(), blob.size(), &rows, &na, &nt, &nb, name, sizeof name, &be_raw));
            HostBytes be(be_raw);
            t_cbor += ms_since(t0);
            lsp_mat* m = nullptr;
            t0 = clk::now();
            CHECK(ctx, lsp_lookup_trace_be(ctx, be.data(), rows, na, nt, nb, publics, &m));
            CHECK(ctx, lsp_ctx_sync(ctx));
            t_witness += ms_since(t0);
            parts.push_back(m);
            // get_air_lookup_config (trace/src/lookup.rs:178-214), shifted by `col`
            LookupIds& L = lids[k];
            for (uint32_t j = 0; j < na; j++) L.a.push_back(col + j);
            for (uint32_t t = 0; t < nt; t++)
                for (uint32_t j = 0; j < nb; j++) L.b.push_back(col + na + t * nb + j);
            uint32_t a_filter = col + na + nt * nb;
            for (uint32_t t = 0; t < nt; t++) L.bf.push_back(a_filter + 1 + t);
            uint32_t a_inv = a_filter + nt + 1;
            for (uint32_t t = 0; t < nt; t++) L.bi.push_back(a_inv + 1 + t);
            for (uint32_t t = 0; t < nt; t++) L.occ.push_back(a_inv + nt + 1 + t);
            lsp_lookup_air_cfg c = {na, L.a.data(), nt, nb, L.b.data(), a_filter, L.bf.data(), a_inv, L.bi.data(), L.occ.data(),
                                    a_inv + 2 * nt + 1};
            lcfg.push_back(c);
            printf("  lookup %s: %zu rows, %u columns, %u tables\n", name, rows, na, nt);
            col += na + nt * (nb + 3) + 3;
            if (height && rows != height) {
                fprintf(stderr, "all sub-traces must have one height (%zu != %zu)\n", rows, height);
                return 1;
            }
            height = rows;
        }
        for (size_t k = 0; k < perms.size(); k++) {
            clk::time_point t0 = clk::now();
            MappedFile blob(perms[k]);
            t_read += ms_since(t0);
            cbor_bytes += blob.size();
            size_t rows = 0;
            uint32_t nc = 0;
            char name[128];
            t0 = clk::now();
            uint8_t* be_raw = nullptr;
            CHECK(ctx, lsp_cbor_permutation_read(blob.data(), blob.size(), &rows, &nc, name, sizeof name, &be_raw));
            HostBytes be(be_raw);
            t_cbor += ms_since(t0);
            lsp_mat* m = nullptr;
            t0 = clk::now();
            CHECK(ctx, lsp_permutation_trace_be(ctx, be.data(), rows, nc, publics, &m));
            CHECK(ctx, lsp_ctx_sync(ctx));
            t_witness += ms_since(t0);
            parts.push_back(m);
            for (uint32_t j = 0; j < nc; j++) {  // trace/src/permutation.rs:84-92, shifted
                pa[k].push_back(col + j);
                pb[k].push_back(col + nc + j);
            }
            lsp_perm_air_cfg c = {nc, pa[k].data(), pb[k].data(), col + 2 * nc, col + 2 * nc + 1};
            pcfg.push_back(c);
            printf("  permutation %s: %zu rows, %u + %u columns\n", name, rows, nc, nc);
            col += 2 * nc + 2;
            if (height && rows != height) {
                fprintf(stderr, "all sub-traces must have one height (%zu != %zu)\n", rows, height);
                return 1;
            }
            height = rows;
        }
        lsp_mat* trace = nullptr;
        const clk::time_point t_cat0 = clk::now();
        CHECK(ctx, lsp_mat_hconcat(ctx, parts.data(), int(parts.size()), &trace));
        CHECK(ctx, lsp_ctx_sync(ctx));
        const double t_cat = ms_since(t_cat0), t_gen = ms_since(t_gen0);
        printf("trace generation [ %.3f ms ]: map files %.3f ms, CBOR parse %.3f ms (%.1f MB, %.0f MB/s), "
               "witness on the device (upload included) %.3f ms, push_traces %.3f ms, other %.3f ms\n",
               t_gen, t_read, t_cbor, cbor_bytes / 1e6, t_cbor > 0 ? cbor_bytes / 1e3 / t_cbor : 0.0, t_witness, t_cat,
               t_gen - t_read - t_cbor - t_witness - t_cat);
        printf("Creating LineaAir...  (%zu rows x %u columns)\n", height, col);

        // ---- prove (main.rs:80-86)
        int log_n = 0;
        while ((size_t(1) << log_n) < height) log_n++;
        int log_q = lsp_air_log_quotient_degree(int(lcfg.size()), int(pcfg.size()));
        size_t words = lsp_proof_words(uint32_t(log_n), col, uint32_t(log_q), &fri);
        std::vector<uint64_t> proof(words);
        float tm[8] = {0};
        printf("Proving...\n");
        CHECK(ctx, lsp_prove_air_dev(ctx, &fri, trace, lcfg.data(), int(lcfg.size()), pcfg.data(), int(pcfg.size()), publics, proof.data(),
                                     words, tm));
        const char* spans[8] = {"commit to trace data: coset_lde_batch", "commit to trace data: merkle tree",
                                "compute quotient polynomial",           "commit to quotient poly chunks",
                                "open: opened values + reduced openings", "FRI prover: commit phase",
                                "FRI prover: grind + query phase",        "proof copy to host"};
        float total = 0;
        for (int i = 0; i < 8; i++) total += tm[i];
        printf("prove [ %.3f ms ]\n", total);
        for (int i = 0; i < 8; i++) printf("  %-44s [ %8.3f ms | %5.1f%% ]\n", spans[i], tm[i], 100.0 * tm[i] / total);
        printf("trace commitment   (Montgomery limbs): %016llx%016llx%016llx%016llx\n", (unsigned long long)proof[3],
               (unsigned long long)proof[2], (unsigned long long)proof[1], (unsigned long long)proof[0]);
        uint64_t h = 1469598103934665603ull;  // FNV-1a over the proof words, for quick comparisons
        for (uint64_t w : proof)
            for (int b = 0; b < 8; b++) h = (h ^ ((w >> (8 * b)) & 0xff)) * 1099511628211ull;
        printf("proof: %zu field elements, fnv1a64 %016llx\n", words / 4, (unsigned long long)h);
        // ---- verify (main.rs:88-96): a fresh transcript over the proof, on the device
        printf("Verifying...\n");
        float verify_ms = 0;
        int verdict = lsp_verify_air(ctx, &fri, uint32_t(log_n), col, lcfg.data(), int(lcfg.size()), pcfg.data(), int(pcfg.size()), publics,
                                     proof.data(), words, &verify_ms);
        if (verdict < 0) CHECK(ctx, verdict);
        printf("verify [ %.3f ms ]: %s\n", verify_ms, verdict == 0 ? "proof accepted" : "PROOF REJECTED");
        if (verdict != 0) {
            fprintf(stderr, "verification failed: LSP_VERIFY code %d\n", verdict);
            return 2;
        }
        if (!out_path.empty()) {
            std::ofstream f(out_path, std::ios::binary);
            f.write(reinterpret_cast<const char*>(proof.data()), std::streamsize(words * 8));
            printf("proof written to %s (layout in DESIGN.md section 7)\n", out_path.c_str());
        }
        lsp_mat_free(ctx, trace);
        for (lsp_mat* m : parts) lsp_mat_free(ctx, m);
        return 0;
    };
    int status = 0;
    for (int rep = 0; rep < repeat && status == 0; rep++) {
        if (repeat > 1) printf("---- pass %d of %d ----\n", rep + 1, repeat);
        status = run_once();
    }
    lsp_ctx_destroy(ctx);
    return status;
}
