// Internal host entry points shared by the C ABI glue and the prover driver.
#pragma once
#include "common.cuh"

namespace lsp {

// GENERATOR = 22 and its inverse, 1/2 (Montgomery form)
__device__ __constant__ const uint32_t FR_GEN[8] = {0xfffffed3u, 0x296c7fffu, 0x6ffffec7u, 0x92921665u,
                                                    0x92860e69u, 0x4c01534du, 0xb9819970u, 0x0c79cfc4u};
__device__ __constant__ const uint32_t FR_GEN_INV[8] = {0xd1745d17u, 0xb76f9745u, 0xafffffffu, 0xfed18274u,
                                                        0x5b36a173u, 0xfce61983u, 0x78dc8d16u, 0x068b6ffdu};
__device__ __constant__ const uint32_t FR_HALF[8] = {0xfffffffau, 0xc396ffffu, 0x1ffffff9u, 0xe6013607u,
                                                     0xd6b1dff7u, 0xbbc63149u, 0x62f41ff9u, 0x0ffb9fc8u};

__device__ __forceinline__ Fr fr_const(const uint32_t* c) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = c[i];
    return r;
}

// TWO_ADIC_ROOT_OF_UNITY = 22^((r-1)/2^47), Montgomery form (ark-bls12-377 FrConfig)
__device__ __constant__ const uint32_t FR_ROOT47[8] = {0xda3ad648u, 0xaf80da4du, 0xfc381dacu, 0x5e223adbu,
                                                       0xb2f92525u, 0x03ba0666u, 0x3befb0ceu, 0x0f906c5bu};

__device__ __forceinline__ Fr fr_two_adic_generator(int bits) {
    Fr w;
#pragma unroll
    for (int i = 0; i < 8; i++) w.l[i] = FR_ROOT47[i];
    for (int i = bits; i < 47; i++) w = fr_sqr(w);
    return w;
}

__device__ __forceinline__ uint32_t bitrev32(uint32_t x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

// ---- ntt.cu ----------------------------------------------------------------
int interpolate_columns(lsp_ctx* ctx, const Fr* in, size_t n, size_t width, Fr* coeffs);
// shift_dev: device pointer to the coset shift (so that data-dependent shifts never visit the host)
int coset_evaluate(lsp_ctx* ctx, const Fr* coeffs, size_t n, size_t width, int added_bits, const Fr* shift_dev, Fr* out);
int coset_evaluate_blocks(lsp_ctx* ctx, const Fr* coeffs, size_t n, size_t width, int added_bits, const Fr* shift_dev, int block0,
                          int n_blocks, Fr* out, size_t out_col_stride);

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

int challenger_init(lsp_ctx* ctx, DevChallenger* ch);
int challenger_observe_dev(lsp_ctx* ctx, DevChallenger* ch, const Fr* vals, int n);
int challenger_observe_host(lsp_ctx* ctx, DevChallenger* ch, const Fr* vals_host, int n);
int challenger_sample(lsp_ctx* ctx, DevChallenger* ch, Fr* out_dev);
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

int quotient_permutation_range(lsp_ctx* ctx, const Fr* lde, size_t lde_rows, size_t p_base, int log_n, int log_q, const PermCfgDev& cfg,
                               const Fr* publics_dev, const Fr* alpha_dev, size_t p0, size_t count, Fr* chunks);

// y[c] = sum_k coeffs[c][k] * z^k
int eval_columns_at(lsp_ctx* ctx, const Fr* coeffs, size_t n, size_t width, const Fr* z_dev, Fr* y_dev);

int fri_fold(lsp_ctx* ctx, const Fr* in, size_t len, const Fr* beta_dev, Fr* out);
int fri_fold_range(lsp_ctx* ctx, const Fr* in, size_t len, size_t j0, size_t h_local, const Fr* beta_dev, Fr* out);

}  // namespace lsp
