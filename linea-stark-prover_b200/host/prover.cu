// Single-GPU entry points of `p3_uni_stark::prove` over `TwoAdicFriPcs` and the standalone pieces (quotient, fold).
// The prover itself lives in host/sharded.cu (one implementation for 1..G GPUs): one stream of kernel launches
// with the transcript kept on the device.
//
// Mirrors, stage for stage (span names from the reference's bench.log:18-67):
//   prove(&config, &air, &mut challenger, trace, &publics)       bin/src/main.rs:80-86
//     "commit to trace data"            pcs.commit -> coset_lde_batch + MerkleTreeMmcs::commit
//     "compute quotient polynomial"     quotient_values with LineaAIR::eval (air/src/lib.rs:116-167)
//     "commit to quotient poly chunks"
//     "open"                            opened values, reduced openings, FRI prover
// The host never reads a challenge: alpha, zeta, the FRI betas and the query
// indices are sampled by single-thread kernels from a device-resident
// HashChallenger and consumed by the next kernels straight from HBM, so the
// whole proof is produced without a host<->device round trip; one D2H copy
// returns it.
//
// Proof layout (flat array of Fr, Montgomery limbs; query indices are raw integers):
//   [0] trace commitment  [1] quotient commitment
//   W trace_local, W trace_next, q quotient-chunk values
//   R commit-phase commitments (R = log2 N - log_final_poly_len)
//   F final-poly coefficients (F = 2^(log_blowup + log_final_poly_len))
//   1 pow witness
//   per query: 1 index | W trace row, log2 L siblings | q chunk row, log2 L siblings |
//              per round r: sibling value, (log2 L - 1 - r) siblings
#include <algorithm>

#include "../csrc/stark.cuh"
#include "comm.hpp"
#include "prove_kernels.cuh"

using namespace lsp;

namespace {

// `get_log_quotient_degree` of p3-uni-stark: `LineaAIR::eval` (air/src/lib.rs:47-54) run on symbolic
// degrees -- trace variable 1, public value 0, constant 0, is_first_row / is_last_row 1, is_transition 0; a product
// adds degrees, a sum or difference takes the larger -- then log2_ceil(max(d_max, 2) - 1).
struct Deg {
    int d;
};
inline Deg operator*(Deg a, Deg b) { return {a.d + b.d}; }
inline Deg operator+(Deg a, Deg b) { return {a.d > b.d ? a.d : b.d}; }
inline Deg operator-(Deg a, Deg b) { return a + b; }

struct SymbolicAir {
    const Deg var{1}, pub{0}, cst{0}, is_first{1}, is_last{1}, is_transition{0};
    int d_max = 0;
    void assert_zero(Deg c) { d_max = c.d > d_max ? c.d : d_max; }
    Deg horner(uint32_t n_cols) const {  // comb = comb * alpha + column, from zero (air/src/lib.rs:129-132)
        Deg acc = cst;
        for (uint32_t i = 0; i < n_cols; i++) acc = acc * pub + var;
        return acc;
    }
    void eval_lookup(const lsp_lookup_air_cfg& c) {  // air/src/lib.rs:57-114
        const Deg a_ch = horner(c.n_a_cols) + pub;
        assert_zero(a_ch * var - cst);                                   // :73
        Deg local_check = var * var, next_check = var * var;             // :75-76
        for (uint32_t t = 0; t < c.n_tables; t++) {
            const Deg b_ch = horner(c.n_b_cols) + pub;
            assert_zero(b_ch * var - cst);                               // :85-88
            local_check = local_check - var * var * var;                 // :90-92
            next_check = next_check - var * var * var;                   // :94-96
        }
        assert_zero(is_first * (var - local_check));                     // :100-102
        assert_zero(is_transition * ((var - var) - next_check));         // :105-107
        assert_zero(is_last * (var - cst));                              // :110-112
    }
    void eval_permutation(const lsp_perm_air_cfg& c) {  // air/src/lib.rs:116-167
        const Deg a_ch = horner(c.n_cols) + pub, b_ch = horner(c.n_cols) + pub;
        assert_zero(b_ch * var - cst);                                   // :143
        assert_zero(is_first * (var - a_ch * var));                      // :146-148
        assert_zero(is_transition * (var - var * a_ch * var));           // :158-161
        assert_zero(is_last * (var - cst));                              // :164-166
    }
};

int log2_ceil(int x) {
    int k = 0;
    while ((1 << k) < x) k++;
    return k;
}

}  // namespace

extern "C" int lsp_air_log_quotient_degree_cfg(const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* perms, int n_perms) {
    if (n_lookups < 0 || n_perms < 0 || (n_lookups && !lookups) || (n_perms && !perms)) return LSP_ERR_PARAM;
    SymbolicAir air;
    for (int i = 0; i < n_lookups; i++) air.eval_lookup(lookups[i]);   // lookups first: trace/src/lib.rs:80-89
    for (int i = 0; i < n_perms; i++) air.eval_permutation(perms[i]);
    return log2_ceil((air.d_max > 2 ? air.d_max : 2) - 1);
}

// The same from the config counts alone, for callers that only know the AIR's shape: every config is taken to have at
// least one column per side and one table (the degree does not depend on the counts beyond that).
extern "C" int lsp_air_log_quotient_degree(int n_lookups, int n_perms) {
    if (n_lookups < 0 || n_perms < 0) return LSP_ERR_PARAM;
    std::vector<lsp_lookup_air_cfg> lk(n_lookups);
    std::vector<lsp_perm_air_cfg> pm(n_perms);
    for (auto& l : lk) {
        memset(&l, 0, sizeof l);
        l.n_a_cols = l.n_tables = l.n_b_cols = 1;
    }
    for (auto& c : pm) {
        memset(&c, 0, sizeof c);
        c.n_cols = 1;
    }
    return lsp_air_log_quotient_degree_cfg(lk.data(), n_lookups, pm.data(), n_perms);
}

extern "C" size_t lsp_proof_words(uint32_t log_n, uint32_t width, uint32_t log_q, const lsp_fri_config* fri) {
    if (!fri || fri->log_final_poly_len >= log_n) return 0;   // zero commit-phase rounds: refused (check_fri_config)
    size_t log_l = size_t(log_n) + fri->log_blowup;
    size_t q = size_t(1) << log_q;
    size_t rounds = log_n - fri->log_final_poly_len;
    size_t f = size_t(1) << (fri->log_blowup + fri->log_final_poly_len);
    size_t per_query = 1 + (width + log_l) + (q + log_l);
    for (size_t r = 0; r < rounds; r++) per_query += 1 + (log_l - 1 - r);
    size_t elems = 2 + 2 * size_t(width) + q + rounds + f + 1 + size_t(fri->num_queries) * per_query;
    return elems * 4;
}

extern "C" int lsp_prove_air_dev(lsp_ctx* ctx, const lsp_fri_config* fri, const lsp_mat* trace, const lsp_lookup_air_cfg* lookups,
                                 int n_lookups, const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                                 uint64_t* proof_out, size_t proof_words, float* timings_ms_out) {
    if (!ctx) return LSP_ERR_PARAM;
    // one rank, no collectives: the sharded prover (host/sharded.cu) IS the prover
    lsp_comm self;
    self.ctx = ctx;
    self.world = 1;
    self.local = true;
    return lsp_prove_air_sharded_dev(&self, fri, trace, lookups, n_lookups, cfgs, n_cfgs, publics, proof_out, proof_words, timings_ms_out);
}

extern "C" int lsp_prove_air(lsp_ctx* ctx, const lsp_fri_config* fri, const uint64_t* trace, size_t rows, size_t width,
                             const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* cfgs, int n_cfgs,
                             const uint64_t publics[2][4], uint64_t* proof_out, size_t proof_words, float* timings_ms_out) {
    if (!ctx || !trace) return LSP_ERR_PARAM;
    lsp_mat* m = nullptr;
    LSP_TRY(lsp_mat_upload(ctx, trace, rows, width, &m));
    int rc = lsp_prove_air_dev(ctx, fri, m, lookups, n_lookups, cfgs, n_cfgs, publics, proof_out, proof_words, timings_ms_out);
    lsp_mat_free(ctx, m);
    return rc;
}

extern "C" int lsp_prove_permutation_dev(lsp_ctx* ctx, const lsp_fri_config* fri, const lsp_mat* trace, const lsp_perm_air_cfg* cfgs,
                                         int n_cfgs, const uint64_t publics[2][4], uint64_t* proof_out, size_t proof_words,
                                         float* timings_ms_out) {
    if (!cfgs || n_cfgs <= 0) return LSP_ERR_PARAM;
    return lsp_prove_air_dev(ctx, fri, trace, nullptr, 0, cfgs, n_cfgs, publics, proof_out, proof_words, timings_ms_out);
}

extern "C" int lsp_prove_permutation(lsp_ctx* ctx, const lsp_fri_config* fri, const uint64_t* trace, size_t rows, size_t width,
                                     const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4], uint64_t* proof_out,
                                     size_t proof_words, float* timings_ms_out) {
    if (!cfgs || n_cfgs <= 0) return LSP_ERR_PARAM;
    return lsp_prove_air(ctx, fri, trace, rows, width, nullptr, 0, cfgs, n_cfgs, publics, proof_out, proof_words, timings_ms_out);
}

// ---------------------------------------------------------------------------
// standalone entry points for the pieces (parity tests / trait-level drop-in)
// ---------------------------------------------------------------------------
extern "C" int lsp_quotient_permutation(lsp_ctx* ctx, const lsp_mat* lde_bitrev, int log_n, int log_q, const lsp_perm_air_cfg* cfgs,
                                        int n_cfgs, const uint64_t publics[2][4], const uint64_t alpha[4], lsp_mat** chunks_out) {
    if (!cfgs || n_cfgs <= 0) return LSP_ERR_PARAM;
    return lsp_quotient_air(ctx, lde_bitrev, log_n, log_q, nullptr, 0, cfgs, n_cfgs, publics, alpha, chunks_out);
}

extern "C" int lsp_quotient_air(lsp_ctx* ctx, const lsp_mat* lde_bitrev, int log_n, int log_q, const lsp_lookup_air_cfg* lookups,
                                int n_lookups, const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                                const uint64_t alpha[4], lsp_mat** chunks_out) {
    if (!ctx || !lde_bitrev || !publics || !alpha || !chunks_out || log_n < 0 || log_q < 0 || log_n + log_q > 31) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch S(ctx);
    PermCfgDev cfg_dev;
    void* blob = nullptr;
    LSP_TRY(upload_air_cfgs(ctx, lookups, n_lookups, cfgs, n_cfgs, lde_bitrev->width, &cfg_dev, &blob));
    S.ptrs.push_back(blob);
    Fr* sc = nullptr;
    LSP_TRY(S.get((void**)&sc, 3 * 32));
    LSP_CUDA(ctx, cudaMemcpyAsync(sc, publics, 64, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemcpyAsync(sc + 2, alpha, 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    lsp_mat* out = nullptr;
    LSP_TRY(mat_alloc(ctx, size_t(1) << log_n, size_t(1) << log_q, &out));
    int rc = quotient_permutation(ctx, lde_bitrev->d, lde_bitrev->rows, log_n, log_q, cfg_dev, sc, sc + 2, out->d);
    if (rc != LSP_OK) {
        lsp_mat_free(ctx, out);
        return rc;
    }
    *chunks_out = out;
    return LSP_OK;
}

extern "C" int lsp_fri_fold(lsp_ctx* ctx, const lsp_mat* in, const uint64_t beta[4], lsp_mat** out) {
    if (!ctx || !in || !beta || !out || in->width != 1 || !is_pow2(in->rows) || in->rows < 2) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch S(ctx);
    Fr* b = nullptr;
    LSP_TRY(S.get((void**)&b, 32));
    LSP_CUDA(ctx, cudaMemcpyAsync(b, beta, 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    lsp_mat* o = nullptr;
    LSP_TRY(mat_alloc(ctx, in->rows / 2, 1, &o));
    int rc = fri_fold(ctx, in->d, in->rows, b, o->d);
    if (rc != LSP_OK) {
        lsp_mat_free(ctx, o);
        return rc;
    }
    *out = o;
    return LSP_OK;
}

// ---------------------------------------------------------------------------
// `Pcs::open` piece by piece, for a trait-level drop-in that keeps its own transcript:
// opened values from the coefficients `lsp_coset_lde_batch` returned, and the reduced-opening vector FRI starts from.
// ---------------------------------------------------------------------------
extern "C" int lsp_eval_at(lsp_ctx* ctx, const lsp_mat* coeffs, const uint64_t z[4], uint64_t* values_out) {
    if (!ctx || !coeffs || !z || !values_out || !is_pow2(coeffs->rows)) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch S(ctx);
    Fr *zd = nullptr, *y = nullptr;
    LSP_TRY(S.get((void**)&zd, 32));
    LSP_TRY(S.get((void**)&y, coeffs->width * 32));
    LSP_CUDA(ctx, cudaMemcpyAsync(zd, z, 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // z[] is caller-owned
    LSP_TRY(eval_columns_at(ctx, coeffs->d, coeffs->rows, coeffs->width, zd, y));
    LSP_CUDA(ctx, cudaMemcpyAsync(values_out, y, coeffs->width * 32, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSP_OK;
}

namespace {
// s[0] = sum_c a^c ys[c]; pw[1] = pw[0] * a^w  (pw[0] = a^offset of this entry)
__global__ void k_entry_scalars(const Fr* __restrict__ alpha, const Fr* __restrict__ ys, int w, Fr* __restrict__ s, Fr* __restrict__ pw) {
    const Fr a = fr_load(alpha);
    Fr acc = fr_load(ys + w - 1);
    for (int i = w - 2; i >= 0; i--) acc = fr_add(fr_mul(acc, a), fr_load(ys + i));
    fr_store(s, acc);
    fr_store(pw + 1, fr_mul(fr_load(pw), fr_pow_u32(a, uint32_t(w))));
}
// out[p] (+)= a^offset * (sum_c a^c m[c][p] - sum_c a^c ys[c]) / (x_p - z)
__global__ void __launch_bounds__(128) k_reduce_entry(const Fr* __restrict__ lde, size_t rows, int w, const Fr* __restrict__ alpha,
                                                      const Fr* __restrict__ s, const Fr* __restrict__ pw, const Fr* __restrict__ inv_den,
                                                      Fr* __restrict__ out, int accumulate) {
    const Fr a = fr_load(alpha), rys = fr_load(s), off = fr_load(pw);
    for (size_t p = blockIdx.x * size_t(blockDim.x) + threadIdx.x; p < rows; p += size_t(gridDim.x) * blockDim.x) {
        Fr rr = fr_load_nc(lde + size_t(w - 1) * rows + p);
        for (int c = w - 2; c >= 0; c--) rr = fr_add(fr_mul(rr, a), fr_load_nc(lde + size_t(c) * rows + p));
        Fr v = fr_mul(fr_mul(off, fr_sub(rr, rys)), fr_load_nc(inv_den + p));
        fr_store(out + p, accumulate ? fr_add(fr_load(out + p), v) : v);
    }
}
}  // namespace

extern "C" int lsp_reduce_openings(lsp_ctx* ctx, const lsp_mat* const* ldes, const uint64_t* points, const uint64_t* const* opened,
                                   int n_entries, const uint64_t alpha[4], lsp_mat** fri_input_out) {
    if (!ctx || !ldes || !points || !opened || n_entries <= 0 || !alpha || !fri_input_out) return LSP_ERR_PARAM;
    const size_t L = ldes[0] ? ldes[0]->rows : 0;
    if (!is_pow2(L) || L < 2) return set_err(ctx, LSP_ERR_PARAM, "LDE height must be a power of two >= 2");
    size_t w_max = 0;
    for (int e = 0; e < n_entries; e++) {
        if (!ldes[e] || !opened[e] || ldes[e]->rows != L) return set_err(ctx, LSP_ERR_PARAM, "all committed matrices share one height on this path");
        w_max = std::max(w_max, ldes[e]->width);
    }
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch S(ctx);
    Fr *a = nullptr, *z = nullptr, *ys = nullptr, *s = nullptr, *pw = nullptr, *inv = nullptr;
    LSP_TRY(S.get((void**)&a, 32));
    LSP_TRY(S.get((void**)&z, size_t(n_entries) * 32));
    LSP_TRY(S.get((void**)&ys, w_max * 32));
    LSP_TRY(S.get((void**)&s, 32));
    LSP_TRY(S.get((void**)&pw, size_t(n_entries + 1) * 32));
    LSP_TRY(S.get((void**)&inv, L * 32));
    LSP_CUDA(ctx, cudaMemcpyAsync(a, alpha, 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemcpyAsync(z, points, size_t(n_entries) * 32, cudaMemcpyHostToDevice, ctx->stream));
    const Fr one = host_pow2_inverse(0);
    LSP_CUDA(ctx, cudaMemcpyAsync(pw, &one, 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    lsp_mat* out = nullptr;
    LSP_TRY(mat_alloc(ctx, L, 1, &out));
    MatGuard guard{ctx, out};
    for (int e = 0; e < n_entries; e++) {  // offset += width per (matrix, point), in the order given (A.9)
        const int w = int(ldes[e]->width);
        LSP_CUDA(ctx, cudaMemcpyAsync(ys, opened[e], size_t(w) * 32, cudaMemcpyHostToDevice, ctx->stream));
        LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        LSP_LAUNCH(ctx, k_entry_scalars, 1, 1, 0, (const Fr*)a, (const Fr*)ys, w, s, pw + e);
        Fr* invp[1] = {inv};
        LSP_TRY(inverse_denominators(ctx, z + e, 1, ilog2(L), invp));
        LSP_LAUNCH(ctx, k_reduce_entry, grid_for(ctx, L, 128), 128, 0, (const Fr*)ldes[e]->d, L, w, (const Fr*)a, (const Fr*)s, (const Fr*)(pw + e),
                   (const Fr*)inv, out->d, e > 0 ? 1 : 0);
    }
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *fri_input_out = guard.release();
    return LSP_OK;
}
