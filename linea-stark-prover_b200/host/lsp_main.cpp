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
//             [--pow-bits 0] [--sbox-d 5] [--device 0] [--gpus 1] [--out proof.bin] [--repeat 1]
//   --gpus N: ONE proof sharded over devices device..device+N-1 (one host thread and one lsp_ctx per GPU, NCCL
//   between them: lsp_prove_air_sharded_dev); every rank builds the witness on its own device from the same parsed bytes.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
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

// One parsed input file: the big-endian bytes of its columns (host) and what they are.
struct SubTrace {
    bool lookup = false;
    std::string name;
    size_t rows = 0;
    uint32_t na = 0, nt = 0, nb = 0;   // lookup: a columns, tables, b columns per table; permutation: na = columns per side
    uint8_t* be = nullptr;
};

// One GPU of the job.
struct Rank {
    int device = 0;
    lsp_ctx* ctx = nullptr;
    lsp_comm* comm = nullptr;
    lsp_mat* trace = nullptr;
};

// body(rank) on one host thread per rank (the library wants one ctx per thread); rank 0 runs on the caller.
template <class F>
void on_all_ranks(std::vector<Rank>& ranks, F body) {
    std::vector<std::thread> pool;
    for (size_t k = 1; k < ranks.size(); k++) pool.emplace_back([&, k]() { body(ranks[k], int(k)); });
    body(ranks[0], 0);
    for (auto& t : pool) t.join();
}

}  // namespace

int main(int argc, char** argv) {
    std::vector<std::string> lookups, perms;
    std::string out_path, out_ser_path;
    uint64_t seed = 0xB200;
    int device = 0, sbox_d = 5, repeat = 1, gpus = 1;
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
        else if (a == "--gpus") gpus = atoi(need("--gpus"));
        else if (a == "--out") out_path = need("--out");
        else if (a == "--out-serialized") out_ser_path = need("--out-serialized");
        else if (a == "--repeat") repeat = atoi(need("--repeat"));
        else {
            fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        }
    }
    if (lookups.empty() && perms.empty()) {
        fprintf(stderr, "usage: lsp_prove [--lookup f.cbor]... [--permutation f.cbor]... [--seed S] [--gpus N] [--out proof.bin]\n");
        return 2;
    }
    if (gpus < 1 || (gpus & (gpus - 1)) || gpus > 8) {  // more ranks than cosets is fine: a rank then owns a fraction of one
        fprintf(stderr, "--gpus must be a power of two, at most 8 (one node)\n");
        return 2;
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

    // ---- one context per GPU; with several, an NCCL communicator between them
    std::vector<Rank> ranks(static_cast<size_t>(gpus));
    uint8_t nccl_id[128] = {0};
    if (gpus > 1 && lsp_nccl_unique_id(nccl_id) != 0) {
        fprintf(stderr, "lsp_nccl_unique_id failed: libnccl.so.2 is needed for --gpus > 1\n");
        return 1;
    }
    on_all_ranks(ranks, [&](Rank& r, int k) {
        r.device = device + k;
        int rc = lsp_ctx_create(r.device, &r.ctx);
        if (rc != 0) {
            fprintf(stderr, "lsp_ctx_create(%d) failed (%d): a CUDA device is required, there is no CPU path\n", r.device, rc);
            exit(1);
        }
        CHECK(r.ctx, lsp_set_poseidon2(r.ctx, 3, sbox_d, 8, 22, consts.data(), &diag[0][0]));
        if (gpus > 1) CHECK(r.ctx, lsp_comm_init_nccl(r.ctx, k, gpus, nccl_id, &r.comm));
    });
    lsp_ctx* ctx = ranks[0].ctx;
    lsp_host_pinned(1);   // parsed columns land in page-locked memory: every rank uploads them at PCIe rate; pinned once, reused by --repeat

    // One pass of main's body.  --repeat N runs it N times in one process: the first pass pays for the CUDA context,
    // the memory pool and the twiddle / selector tables; later passes are what a long-running prover sees.
    auto run_once = [&]() -> int {
        using clk = std::chrono::steady_clock;
        auto ms_since = [](clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); };
        // ---- read_file (main.rs:37-41): parse every input once on the host
        printf("Generating trace...\n");
        const clk::time_point t_gen0 = clk::now();
        double t_read = 0, t_cbor = 0;
        size_t cbor_bytes = 0;
        std::vector<SubTrace> subs;
        // `push_traces` (trace/src/lib.rs:62-79): the trace height is the tallest column of ANY input file, and every
        // sub-trace is resized to it (zero rows) before its witness is built.  With one file its own height is that
        // height; with several a shape pass over each finds the maximum first.
        size_t max_height = 0;
        if (lookups.size() + perms.size() > 1) {
            auto shape = [&](const std::string& path, bool is_lookup) {
                MappedFile blob(path);
                size_t rows = 0;
                uint32_t a = 0, t = 0, b = 0;
                if (is_lookup)
                    CHECK(ctx, lsp_cbor_lookup_shape(blob.data(), blob.size(), &rows, &a, &t, &b, nullptr, 0));
                else
                    CHECK(ctx, lsp_cbor_permutation_shape(blob.data(), blob.size(), &rows, &a, nullptr, 0));
                max_height = std::max(max_height, rows);
            };
            for (auto& f : lookups) shape(f, true);
            for (auto& f : perms) shape(f, false);
        }
        auto load = [&](const std::string& path, bool is_lookup) {
            clk::time_point t0 = clk::now();
            MappedFile blob(path);
            t_read += ms_since(t0);
            cbor_bytes += blob.size();
            SubTrace st;
            st.lookup = is_lookup;
            char name[128] = {0};
            t0 = clk::now();
            if (is_lookup)
                CHECK(ctx, lsp_cbor_lookup_read_rows(blob.data(), blob.size(), max_height, &st.rows, &st.na, &st.nt, &st.nb, name, sizeof name, &st.be));
            else
                CHECK(ctx, lsp_cbor_permutation_read_rows(blob.data(), blob.size(), max_height, &st.rows, &st.na, name, sizeof name, &st.be));
            t_cbor += ms_since(t0);
            st.name = name;
            subs.push_back(st);
        };
        for (auto& f : lookups) load(f, true);      // push_traces: lookups first, then permutations (trace/src/lib.rs:62-92)
        for (auto& f : perms) load(f, false);
        struct FreeSubs {
            std::vector<SubTrace>& v;
            ~FreeSubs() {
                for (auto& st : v) lsp_host_free(st.be);
            }
        } free_subs{subs};

        // ---- AIR configs, shifted by the running width exactly as `cfg.shift(self.columns.len())` does
        std::vector<lsp_lookup_air_cfg> lcfg;
        std::vector<lsp_perm_air_cfg> pcfg;
        std::vector<LookupIds> lids(lookups.size());
        std::vector<std::vector<uint32_t>> pa(perms.size()), pb(perms.size());
        uint32_t col = 0;
        size_t height = 0, li = 0, pi = 0;
        for (const SubTrace& st : subs) {
            if (height && st.rows != height) {  // cannot happen: every file was read at the common height
                fprintf(stderr, "internal error: sub-trace heights differ (%zu != %zu)\n", st.rows, height);
                return 1;
            }
            height = st.rows;
            if (st.lookup) {
                const uint32_t na = st.na, nt = st.nt, nb = st.nb;
                LookupIds& L = lids[li++];  // get_air_lookup_config (trace/src/lookup.rs:178-214), shifted by `col`
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
                printf("  lookup %s: %zu rows, %u columns, %u tables\n", st.name.c_str(), st.rows, na, nt);
                col += na + nt * (nb + 3) + 3;
            } else {
                const uint32_t nc = st.na;
                for (uint32_t j = 0; j < nc; j++) {  // trace/src/permutation.rs:84-92, shifted
                    pa[pi].push_back(col + j);
                    pb[pi].push_back(col + nc + j);
                }
                lsp_perm_air_cfg c = {nc, pa[pi].data(), pb[pi].data(), col + 2 * nc, col + 2 * nc + 1};
                pcfg.push_back(c);
                pi++;
                printf("  permutation %s: %zu rows, %u + %u columns\n", st.name.c_str(), st.rows, nc, nc);
                col += 2 * nc + 2;
            }
        }

        // ---- get_trace + push_traces on every GPU (each rank needs the whole trace to interpolate its columns)
        const clk::time_point t_wit0 = clk::now();
        on_all_ranks(ranks, [&](Rank& r, int) {
            std::vector<lsp_mat*> parts;
            for (const SubTrace& st : subs) {
                lsp_mat* m = nullptr;
                if (st.lookup)
                    CHECK(r.ctx, lsp_lookup_trace_be(r.ctx, st.be, st.rows, st.na, st.nt, st.nb, publics, &m));
                else
                    CHECK(r.ctx, lsp_permutation_trace_be(r.ctx, st.be, st.rows, st.na, publics, &m));
                parts.push_back(m);
            }
            CHECK(r.ctx, lsp_mat_hconcat(r.ctx, parts.data(), int(parts.size()), &r.trace));
            CHECK(r.ctx, lsp_ctx_sync(r.ctx));
            for (lsp_mat* m : parts) lsp_mat_free(r.ctx, m);
        });
        const double t_witness = ms_since(t_wit0), t_gen = ms_since(t_gen0);
        printf("trace generation [ %.3f ms ]: map files %.3f ms, CBOR parse %.3f ms (%.1f MB, %.0f MB/s), "
               "witness on the device%s (upload and push_traces included) %.3f ms, other %.3f ms\n",
               t_gen, t_read, t_cbor, cbor_bytes / 1e6, t_cbor > 0 ? cbor_bytes / 1e3 / t_cbor : 0.0, gpus > 1 ? "s" : "", t_witness,
               t_gen - t_read - t_cbor - t_witness);
        printf("Creating LineaAir...  (%zu rows x %u columns)\n", height, col);

        // ---- prove (main.rs:80-86)
        int log_n = 0;
        while ((size_t(1) << log_n) < height) log_n++;
        int log_q = lsp_air_log_quotient_degree_cfg(lcfg.data(), int(lcfg.size()), pcfg.data(), int(pcfg.size()));
        size_t words = lsp_proof_words(uint32_t(log_n), col, uint32_t(log_q), &fri);
        std::vector<std::vector<uint64_t>> proofs(ranks.size(), std::vector<uint64_t>(words));
        std::vector<uint64_t>& proof = proofs[0];
        float tm[8] = {0};
        printf("Proving...%s\n", gpus > 1 ? "  (one proof sharded by LDE row ranges)" : "");
        const clk::time_point t_prove0 = clk::now();
        on_all_ranks(ranks, [&](Rank& r, int k) {
            float tmk[8] = {0};
            if (gpus == 1)
                CHECK(r.ctx, lsp_prove_air_dev(r.ctx, &fri, r.trace, lcfg.data(), int(lcfg.size()), pcfg.data(), int(pcfg.size()), publics,
                                               proofs[size_t(k)].data(), words, tmk));
            else
                CHECK(r.ctx, lsp_prove_air_sharded_dev(r.comm, &fri, r.trace, lcfg.data(), int(lcfg.size()), pcfg.data(), int(pcfg.size()),
                                                       publics, proofs[size_t(k)].data(), words, tmk));
            if (k == 0) memcpy(tm, tmk, sizeof tm);
        });
        const double t_prove_wall = ms_since(t_prove0);
        for (size_t k = 1; k < proofs.size(); k++)
            if (proofs[k] != proof) {
                fprintf(stderr, "rank %zu holds a different proof than rank 0\n", k);
                return 1;
            }
        const char* spans[8] = {"commit to trace data: coset_lde_batch", "commit to trace data: merkle tree",
                                "compute quotient polynomial",           "commit to quotient poly chunks",
                                "open: opened values + reduced openings", "FRI prover: commit phase",
                                "FRI prover: grind + query phase",        "proof copy to host"};
        float total = 0;
        for (int i = 0; i < 8; i++) total += tm[i];
        printf("prove [ %.3f ms device, %.3f ms wall%s ]\n", total, t_prove_wall, gpus > 1 ? ", all ranks" : "");
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
        for (Rank& r : ranks) {
            lsp_mat_free(r.ctx, r.trace);
            r.trace = nullptr;
        }
        if (verdict != 0) {
            fprintf(stderr, "verification failed: LSP_VERIFY code %d\n", verdict);
            return 2;
        }
        if (!out_path.empty()) {
            std::ofstream f(out_path, std::ios::binary);
            f.write(reinterpret_cast<const char*>(proof.data()), std::streamsize(words * 8));
            printf("proof written to %s (layout in DESIGN.md section 7)\n", out_path.c_str());
        }
        if (!out_ser_path.empty()) {  // `Proof` in struct order (lsp_proof_serialize): what a verifier elsewhere would receive
            std::vector<uint8_t> bytes(lsp_proof_serialized_bytes(uint32_t(log_n), col, uint32_t(log_q), &fri));
            size_t n_bytes = 0;
            CHECK(ctx, lsp_proof_serialize(proof.data(), words, uint32_t(log_n), col, uint32_t(log_q), &fri, bytes.data(), bytes.size(), &n_bytes));
            std::ofstream f(out_ser_path, std::ios::binary);
            f.write(reinterpret_cast<const char*>(bytes.data()), std::streamsize(n_bytes));
            printf("serialised proof (%zu bytes) written to %s\n", n_bytes, out_ser_path.c_str());
        }
        return 0;
    };
    int status = 0;
    for (int rep = 0; rep < repeat && status == 0; rep++) {
        if (repeat > 1) printf("---- pass %d of %d ----\n", rep + 1, repeat);
        status = run_once();
    }
    lsp_host_pinned(0);
    on_all_ranks(ranks, [&](Rank& r, int) {
        if (r.comm) lsp_comm_destroy(r.comm);
        lsp_ctx_destroy(r.ctx);
    });
    return status;
}
