// Internal host entry points shared by the C ABI glue and the prover driver.
#pragma once
#include "common.cuh"

namespace lsp {

// 1/2 (Montgomery form)
__device__ __constant__ const uint32_t FR_HALF[8] = {0xfffffffau, 0xc396ffffu, 0x1ffffff9u, 0xe6013607u,
                                                     0xd6b1dff7u, 0xbbc63149u, 0x62f41ff9u, 0x0ffb9fc8u};

__device__ __forceinline__ Fr fr_const(const uint32_t* c) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = c[i];
    return r;
}

// `Bls12_377Fr::GENERATOR` (the coset shift of every committed LDE) and `two_adic_generator(47)`: the fork-only crate that
// fixes them is unavailable, so they are run-time parameters of the context (lsp_set_field_consts),
// resident in device memory; the defaults are arkworks' FrConfig values (GENERATOR = 22, TWO_ADIC_ROOT_OF_UNITY).
// (FieldConsts itself is declared in common.cuh.)
constexpr uint32_t FR_GEN_DEFAULT[8] = {0xfffffed3u, 0x296c7fffu, 0x6ffffec7u, 0x92921665u, 0x92860e69u, 0x4c01534du, 0xb9819970u, 0x0c79cfc4u};
constexpr uint32_t FR_ROOT47_DEFAULT[8] = {0xda3ad648u, 0xaf80da4du, 0xfc381dacu, 0x5e223adbu, 0xb2f92525u, 0x03ba0666u, 0x3befb0ceu, 0x0f906c5bu};

__device__ __forceinline__ Fr fr_two_adic_generator(const FieldConsts* fc, int bits) {
    Fr w = fr_load(&fc->root47);
    for (int i = bits; i < 47; i++) w = fr_sqr(w);
    return w;
}

__device__ __forceinline__ uint32_t bitrev32(uint32_t x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

// 2^-k in Montgomery form, on the host: halve (R mod r) k times with 4 x u64 arithmetic.  (1/N of the inverse
// transform, 1/F of the final-polynomial interpolation: kernel arguments, never read back from the device.)
inline Fr host_pow2_inverse(int k) {
    uint64_t v[4] = {0x7d1c7ffffffffff3ull, 0x7257f50f6ffffff2ull, 0x16d81575512c0feeull, 0x0d4bda322bbb9a9dull};
    static const uint64_t Pm[4] = {0x0a11800000000001ull, 0x59aa76fed0000001ull, 0x60b44d1e5c37b001ull, 0x12ab655e9a2ca556ull};
    for (int i = 0; i < k; i++) {
        if (v[0] & 1) {
            unsigned __int128 c = 0;
            for (int j = 0; j < 4; j++) {
                c += (unsigned __int128)v[j] + Pm[j];
                v[j] = (uint64_t)c;
                c >>= 64;
            }
        }
        for (int j = 0; j < 3; j++) v[j] = (v[j] >> 1) | (v[j + 1] << 63);
        v[3] >>= 1;
    }
    Fr r;
    memcpy(r.l, v, 32);
    return r;
}

// ---- ntt.cu ----------------------------------------------------------------
int interpolate_columns(lsp_ctx* ctx, const Fr* in, size_t n, size_t width, Fr* coeffs);
// shift_dev: device pointer to the coset shift (so that data-dependent shifts never visit the host)
int coset_evaluate(lsp_ctx* ctx, const Fr* coeffs, size_t n, size_t width, int added_bits, const Fr* shift_dev, Fr* out);
int coset_evaluate_blocks(lsp_ctx* ctx, const Fr* coeffs, size_t n, size_t width, int added_bits, const Fr* shift_dev, int block0,
                          int n_blocks, Fr* out, size_t out_col_stride);
// rows [sub*M, (sub+1)*M), M = n >> log_s, of row block `block` only (a rank that owns a fraction of a coset)
int coset_evaluate_subblock(lsp_ctx* ctx, const Fr* coeffs, size_t n, size_t width, int added_bits, const Fr* shift_dev, int block, int log_s,
                            int sub, Fr* out, size_t out_col_stride);

// ---- core.cu ---------------------------------------------------------------
// device row-major (host layout) -> device column-major
int rowmajor_to_colmajor(lsp_ctx* ctx, const Fr* rm_dev, size_t rows, size_t width, Fr* cm_dev);
int merkle_build(lsp_ctx* ctx, const Fr* const* d_cols, int width, size_t h, Fr* digests);
// digest layers over a vector viewed as rows of 2 (FRI commit-phase trees): layer0[j] = hash([v[2j], v[2j+1]])
int merkle_build_pairs(lsp_ctx* ctx, const Fr* vec, size_t len, Fr* digests);

// ---- stark.cu --------------------------------------------------------------
constexpr int CH_CAP = 2048;
struct DevChallenger {  // HashChallenger<Val,Hash,1> state, resident in device memory
    Fr input[CH_CAP];
    int n_input;
    int overflow;
};

// `FriConfig` (bin/src/main.rs:58-64) against a trace of 2^log_n rows and 2^log_q quotient chunks: the one place the
// prover, the sharded prover and the verifier check it.
//  * log_final_poly_len >= log_n would leave ZERO commit-phase rounds: the reduced opening only enters the fold chain
//    inside a round, so the low-degree test would check nothing about the committed data (the pinned verifier then
//    rejects every honest proof with FinalPolyMismatch and accepts a forged all-zero final polynomial).  Refused.
//  * proof_of_work_bits > 32: `sample_bits` reads the low limb of the canonical integer; more bits than that would
//    silently be enforced as 32.  Refused.
inline int check_fri_config(lsp_ctx* ctx, const lsp_fri_config* fri, int log_n, int log_q) {
    const int log_b = int(fri->log_blowup), log_l = log_n + log_b;
    if (log_q > log_b) return set_err(ctx, LSP_ERR_PARAM, "quotient degree 2^%d exceeds blowup 2^%d", log_q, log_b);
    if (log_l > 31 || log_l < 1) return set_err(ctx, LSP_ERR_PARAM, "LDE of 2^%d rows unsupported", log_l);
    if (int(fri->log_final_poly_len) >= log_n)
        return set_err(ctx, LSP_ERR_PARAM, "log_final_poly_len %u leaves no commit-phase round for a trace of 2^%d rows (vacuous low-degree test)",
                       fri->log_final_poly_len, log_n);
    if (fri->num_queries == 0 || fri->num_queries > 4096) return set_err(ctx, LSP_ERR_PARAM, "num_queries out of range");
    if (fri->proof_of_work_bits > 32) return set_err(ctx, LSP_ERR_PARAM, "proof_of_work_bits %u > 32 unsupported", fri->proof_of_work_bits);
    if ((size_t(1) << (log_b + int(fri->log_final_poly_len))) + 8 > size_t(CH_CAP))
        return set_err(ctx, LSP_ERR_PARAM, "final polynomial of 2^%d coefficients unsupported", log_b + int(fri->log_final_poly_len));
    return LSP_OK;
}

struct PermCfgDev {  // flattened LineaAIR config list in device memory: lookups first, then permutations
    // lookup i lives at lk + lk_off[i]:
    //   [n_a, n_tables, n_b, a_filter, a_inverses, check, a_ids[n_a], then per table: b_filter, b_inverses, occurrences, b_ids[n_b]]
    int n_lookups = 0;
    const uint32_t* lk_off = nullptr;
    const uint32_t* lk = nullptr;
    int n_cfgs;
    const uint32_t* n_cols;    // per cfg
    const uint32_t* ids_off;   // per cfg, offset into ids (a ids then b ids)
    const uint32_t* ids;
    const uint32_t* b_inverse_id;
    const uint32_t* check_id;
};

// LineaAIR::eval folded by powers of the STARK challenge (ProverConstraintFolder / VerifierConstraintFolder,
// acc = acc * alpha + constraint): lookups first, then permutations, as `oracle/air.py` orders them.
// Row `p` is the local row, `pn` the next one; element (column c, row r) is lde[c * lde_rows + r].  The quotient
// kernel calls it on LDE rows, the verifier on the opened values (lde_rows = 1, p = 0, pn = width).  `pn` is a signed
// element offset: a caller whose next rows live in a second matrix of the same shape passes pn = row + (next - lde).
__device__ __forceinline__ Fr fold_air_constraints(const PermCfgDev& cfg, const Fr* __restrict__ lde, size_t lde_rows, long long p,
                                                   long long pn, const Fr& alpha_air, const Fr& delta, const Fr& alpha,
                                                   const Fr& is_first, const Fr& is_last, const Fr& is_trans) {
    const Fr one = fr_one();
    Fr acc = fr_zero();
    bool first_c = true;
    // ---- eval_lookup (air/src/lib.rs:57-114), one pass per AirLookupConfig ----------------
    for (int k = 0; k < cfg.n_lookups; k++) {
        const uint32_t* r = cfg.lk + cfg.lk_off[k];
        const uint32_t n_a = r[0], n_t = r[1], n_b = r[2];
        const Fr* col_af = lde + size_t(r[3]) * lde_rows;
        const Fr* col_ai = lde + size_t(r[4]) * lde_rows;
        const Fr* col_chk = lde + size_t(r[5]) * lde_rows;
        const uint32_t* a_ids = r + 6;
        Fr a_l = fr_load_nc(lde + size_t(a_ids[0]) * lde_rows + p);        // :65-68
        for (uint32_t j = 1; j < n_a; j++) a_l = fr_add(fr_mul(a_l, alpha_air), fr_load_nc(lde + size_t(a_ids[j]) * lde_rows + p));
        a_l = fr_add(a_l, delta);                                               // :70
        const Fr ai_l = fr_load_nc(col_ai + p), ai_n = fr_load_nc(col_ai + pn);
        Fr c = fr_sub(fr_mul(a_l, ai_l), one);                                  // :73
        acc = first_c ? c : fr_add(fr_mul(acc, alpha), c);
        first_c = false;
        Fr chk_l_expr = fr_mul(fr_load_nc(col_af + p), ai_l);                   // :75
        Fr chk_n_expr = fr_mul(fr_load_nc(col_af + pn), ai_n);                  // :76
        const uint32_t* t_rec = a_ids + n_a;
        for (uint32_t t = 0; t < n_t; t++, t_rec += 3 + n_b) {
            const Fr* col_bf = lde + size_t(t_rec[0]) * lde_rows;
            const Fr* col_bi = lde + size_t(t_rec[1]) * lde_rows;
            const Fr* col_oc = lde + size_t(t_rec[2]) * lde_rows;
            const uint32_t* b_ids = t_rec + 3;
            Fr b_l = fr_load_nc(lde + size_t(b_ids[0]) * lde_rows + p);     // :79-82
            for (uint32_t j = 1; j < n_b; j++) b_l = fr_add(fr_mul(b_l, alpha_air), fr_load_nc(lde + size_t(b_ids[j]) * lde_rows + p));
            b_l = fr_add(b_l, delta);                                           // :84
            const Fr bi_l = fr_load_nc(col_bi + p), bi_n = fr_load_nc(col_bi + pn);
            c = fr_sub(fr_mul(b_l, bi_l), one);                                 // :85-88
            acc = fr_add(fr_mul(acc, alpha), c);
            chk_l_expr = fr_sub(chk_l_expr, fr_mul(fr_mul(fr_load_nc(col_bf + p), fr_load_nc(col_oc + p)), bi_l));     // :90-92
            chk_n_expr = fr_sub(chk_n_expr, fr_mul(fr_mul(fr_load_nc(col_bf + pn), fr_load_nc(col_oc + pn)), bi_n));   // :94-96
        }
        const Fr chk_l = fr_load_nc(col_chk + p), chk_n = fr_load_nc(col_chk + pn);
        acc = fr_add(fr_mul(acc, alpha), fr_mul(is_first, fr_sub(chk_l, chk_l_expr)));                  // :100-102
        acc = fr_add(fr_mul(acc, alpha), fr_mul(is_trans, fr_sub(fr_sub(chk_n, chk_l), chk_n_expr)));  // :105-107
        acc = fr_add(fr_mul(acc, alpha), fr_mul(is_last, chk_l));                                       // :110-112
    }
    // ---- eval_permutation (air/src/lib.rs:116-167) -----------------------------------------
    for (int k = 0; k < cfg.n_cfgs; k++) {
        const uint32_t nc = cfg.n_cols[k];
        const uint32_t* a_ids = cfg.ids + cfg.ids_off[k];
        const uint32_t* b_ids = a_ids + nc;
        const Fr* col_inv = lde + size_t(cfg.b_inverse_id[k]) * lde_rows;
        const Fr* col_chk = lde + size_t(cfg.check_id[k]) * lde_rows;
        // Horner combinations (air/src/lib.rs:129-137,150-153): comb = comb*alpha + col
        Fr a_l = fr_load_nc(lde + size_t(a_ids[0]) * lde_rows + p);
        Fr b_l = fr_load_nc(lde + size_t(b_ids[0]) * lde_rows + p);
        Fr a_n = fr_load_nc(lde + size_t(a_ids[0]) * lde_rows + pn);
        for (uint32_t j = 1; j < nc; j++) {
            a_l = fr_add(fr_mul(a_l, alpha_air), fr_load_nc(lde + size_t(a_ids[j]) * lde_rows + p));
            b_l = fr_add(fr_mul(b_l, alpha_air), fr_load_nc(lde + size_t(b_ids[j]) * lde_rows + p));
            a_n = fr_add(fr_mul(a_n, alpha_air), fr_load_nc(lde + size_t(a_ids[j]) * lde_rows + pn));
        }
        a_l = fr_add(a_l, delta);
        b_l = fr_add(b_l, delta);
        a_n = fr_add(a_n, delta);
        Fr inv_l = fr_load_nc(col_inv + p), inv_n = fr_load_nc(col_inv + pn);
        Fr chk_l = fr_load_nc(col_chk + p), chk_n = fr_load_nc(col_chk + pn);
        // C0: b_ch * inv - 1                                  (:143)
        Fr c0 = fr_sub(fr_mul(b_l, inv_l), one);
        // C1: is_first * (check - a_ch * inv)                  (:146-148)
        Fr c1 = fr_mul(is_first, fr_sub(chk_l, fr_mul(a_l, inv_l)));
        // C2: is_transition * (check' - check * a_ch' * inv')  (:158-161)
        Fr c2 = fr_mul(is_trans, fr_sub(chk_n, fr_mul(fr_mul(chk_l, a_n), inv_n)));
        // C3: is_last * (check - 1)                            (:164-166)
        Fr c3 = fr_mul(is_last, fr_sub(chk_l, one));
        acc = first_c ? c0 : fr_add(fr_mul(acc, alpha), c0);
        first_c = false;
        acc = fr_add(fr_mul(acc, alpha), c1);
        acc = fr_add(fr_mul(acc, alpha), c2);
        acc = fr_add(fr_mul(acc, alpha), c3);
    }
    return acc;
}

// Everything the verifier's transcript reads and writes (device pointers; the proof regions point into the
// uploaded proof).  scal receives [alpha, zeta, alpha_fri].
enum { VT_ALPHA = 0, VT_ZETA = 1, VT_ALPHA_FRI = 2, VT_COUNT = 3 };
struct VerifyTranscriptArgs {
    int log_n, n_rounds, n_final, pow_bits, log_l, n_queries;
    int alpha_before_openings, observe_opened_values, n_opened;   // lsp_set_transcript_flags; opened values: 2W + q
    const Fr *trace_commit, *quot_commit, *publics, *opened, *fri_commits, *final_poly, *pow_witness;
    Fr *scal, *betas;
    uint32_t *pow_low, *idx;
};
int verify_transcript(lsp_ctx* ctx, DevChallenger* ch, const VerifyTranscriptArgs& A);

int challenger_init(lsp_ctx* ctx, DevChallenger* ch);
int challenger_observe_dev(lsp_ctx* ctx, DevChallenger* ch, const Fr* vals, int n);
int challenger_observe_host(lsp_ctx* ctx, DevChallenger* ch, const Fr* vals_host, int n);
int challenger_sample(lsp_ctx* ctx, DevChallenger* ch, Fr* out_dev);
// observe(vals[0..n)) (copied to `copy_to` too when non-null), then sample(): one launch.  The list form observes several
// runs of values in order (the first run is the one copied).
struct ObserveList {
    const Fr* p[4];
    int n[4];
    int n_seg;
};
int challenger_observe_sample(lsp_ctx* ctx, DevChallenger* ch, const Fr* vals, int n, Fr* copy_to, Fr* out_dev);
int challenger_observe_sample(lsp_ctx* ctx, DevChallenger* ch, const ObserveList& L, Fr* copy_to, Fr* out_dev);
// `n` successive sample_bits(bits) -> idx_out[n]
int challenger_sample_bits(lsp_ctx* ctx, DevChallenger* ch, int bits, int n, uint32_t* idx_out);
// grind(bits): smallest witness; observes it and consumes one sample (check_witness)
int challenger_grind(lsp_ctx* ctx, DevChallenger* ch, int bits, Fr* witness_out);

int upload_perm_cfgs(lsp_ctx* ctx, const lsp_perm_air_cfg* cfgs, int n_cfgs, size_t width, PermCfgDev* out, void** blob);
// lookups (may be empty) followed by permutations (may be empty); checks ids against `width` and that the widths add up to it
int upload_air_cfgs(lsp_ctx* ctx, const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* perms, int n_perms,
                    size_t width, PermCfgDev* out, void** blob);

// E[p] = 1/(g*w_L^{bitrev(p)} - z) for p < 2^log_m, one array per point (points in device memory)
int inverse_denominators(lsp_ctx* ctx, const Fr* z_dev, int n_points, int log_m, Fr* const* out);
// same for rows [p0, p0+count) only (out[p] holds `count` entries)
int inverse_denominators_range(lsp_ctx* ctx, const Fr* z_dev, int n_points, int log_m, size_t p0, size_t count, Fr* const* out);

int quotient_permutation(lsp_ctx* ctx, const Fr* lde, size_t lde_rows, int log_n, int log_q, const PermCfgDev& cfg,
                         const Fr* publics_dev, const Fr* alpha_dev, Fr* chunks /* q columns of N */);

// storage rows [p0, p0 + count) of the quotient domain; `lde` points at storage row p_base.  lde_next == nullptr: the next row
// of every row lies in the same matrix (whole cosets).  Else the range is a sub-coset of `count` points (a rank that owns a
// fraction of a coset) and the next rows are those of the neighbouring sub-coset, held in lde_next (same shape): at the same
// local row, or -- next_rotated, the neighbour being the residue that wrapped to 0 -- at the row of the following point.
int quotient_permutation_range(lsp_ctx* ctx, const Fr* lde, size_t lde_rows, size_t p_base, int log_n, int log_q, const PermCfgDev& cfg,
                               const Fr* publics_dev, const Fr* alpha_dev, size_t p0, size_t count, Fr* chunks, const Fr* lde_next = nullptr,
                               bool next_rotated = false);

// y[c] = sum_k coeffs[c][k] * z^k
int eval_columns_at(lsp_ctx* ctx, const Fr* coeffs, size_t n, size_t width, const Fr* z_dev, Fr* y_dev);

int fri_fold(lsp_ctx* ctx, const Fr* in, size_t len, const Fr* beta_dev, Fr* out);
int fri_fold_range(lsp_ctx* ctx, const Fr* in, size_t len, size_t j0, size_t h_local, const Fr* beta_dev, Fr* out);

}  // namespace lsp
