// Width-3 Poseidon2 over BLS12-377 Fr, state held in registers (24 x u32).
//
// Device-side replacement for the reference's
//   Perm     = Poseidon2Bls12337<3>                      (bin/src/config.rs:11, bin/src/main.rs:49)
//   Hash     = PaddingFreeSponge<Perm,3,2,1>             (bin/src/config.rs:12)
//   Compress = CompressionFunctionFromHasher<Hash,2,1>   (bin/src/config.rs:17)
// Round constants, round counts, S-box degree and the internal diagonal are
// runtime parameters (lsp_set_poseidon2): the reference draws the constants
// at start-up (bin/src/main.rs:49) and the fork-only crate that fixes the rest
// is not available, so nothing is assumed.
#pragma once
#include "fr.cuh"

namespace lsp {

constexpr int P2_WIDTH = 3;
constexpr int P2_MAX_HALF_F = 8;
constexpr int P2_MAX_P = 64;

// Passed by value as a __grid_constant__ kernel parameter (constant bank).
struct P2Params {
    int half_f;    // rounds_f / 2
    int rounds_p;
    int sbox_d;
    int diag_kind;  // 0: generic diag, 1: (1,1,2)
    Fr ext_initial[P2_MAX_HALF_F][3];
    Fr ext_terminal[P2_MAX_HALF_F][3];
    Fr internal[P2_MAX_P];
    Fr diag_m1[3];
};

template <int D>
__device__ __forceinline__ Fr p2_sbox(const Fr& x) {
    if (D == 3) {
        return fr_mul_call(fr_sqr_call(x), x);
    } else if (D == 5) {
        Fr x2 = fr_sqr_call(x);
        return fr_mul_call(fr_sqr_call(x2), x);
    } else if (D == 7) {
        Fr x2 = fr_sqr_call(x);
        Fr x4 = fr_sqr_call(x2);
        return fr_mul_call(fr_mul_call(x4, x2), x);
    } else if (D == 11) {
        Fr x2 = fr_sqr_call(x);
        Fr x8 = fr_sqr_call(fr_sqr_call(x2));
        return fr_mul_call(fr_mul_call(x8, x2), x);
    } else {  // 17
        Fr x16 = fr_sqr_call(fr_sqr_call(fr_sqr_call(fr_sqr_call(x))));
        return fr_mul_call(x16, x);
    }
}

__device__ __forceinline__ void p2_ext_linear(Fr& s0, Fr& s1, Fr& s2) {
    Fr t = fr_add(fr_add(s0, s1), s2);
    s0 = fr_add(s0, t);
    s1 = fr_add(s1, t);
    s2 = fr_add(s2, t);
}

// Not inlined: one permutation is ~25k instructions, so the call is free and every
// kernel of a translation unit shares one body per S-box degree (I-cache, build time).
template <int D>
__device__ __noinline__ void p2_permute(const P2Params& P, Fr& s0, Fr& s1, Fr& s2) {
    p2_ext_linear(s0, s1, s2);
#pragma unroll 1
    for (int r = 0; r < P.half_f; r++) {
        s0 = p2_sbox<D>(fr_add(s0, P.ext_initial[r][0]));
        s1 = p2_sbox<D>(fr_add(s1, P.ext_initial[r][1]));
        s2 = p2_sbox<D>(fr_add(s2, P.ext_initial[r][2]));
        p2_ext_linear(s0, s1, s2);
    }
    if (P.diag_kind == 1) {
#pragma unroll 1
        for (int r = 0; r < P.rounds_p; r++) {
            s0 = p2_sbox<D>(fr_add(s0, P.internal[r]));
            Fr t = fr_add(fr_add(s0, s1), s2);
            s0 = fr_add(s0, t);
            s1 = fr_add(s1, t);
            s2 = fr_add(fr_dbl(s2), t);
        }
    } else {
#pragma unroll 1
        for (int r = 0; r < P.rounds_p; r++) {
            s0 = p2_sbox<D>(fr_add(s0, P.internal[r]));
            Fr t = fr_add(fr_add(s0, s1), s2);
            s0 = fr_add(fr_mul_call(s0, P.diag_m1[0]), t);
            s1 = fr_add(fr_mul_call(s1, P.diag_m1[1]), t);
            s2 = fr_add(fr_mul_call(s2, P.diag_m1[2]), t);
        }
    }
#pragma unroll 1
    for (int r = 0; r < P.half_f; r++) {
        s0 = p2_sbox<D>(fr_add(s0, P.ext_terminal[r][0]));
        s1 = p2_sbox<D>(fr_add(s1, P.ext_terminal[r][1]));
        s2 = p2_sbox<D>(fr_add(s2, P.ext_terminal[r][2]));
        p2_ext_linear(s0, s1, s2);
    }
}

// compress(l, r) = perm([l, r, 0])[0]
template <int D>
__device__ __forceinline__ Fr p2_compress(const P2Params& P, const Fr& l, const Fr& r) {
    Fr s0 = l, s1 = r, s2 = fr_zero();
    p2_permute<D>(P, s0, s1, s2);
    return s0;
}

// Dispatch a templated launch on the runtime S-box degree.
#define LSP_DISPATCH_SBOX(d, ...)                  \
    switch (d) {                                   \
        case 3: { constexpr int D = 3; __VA_ARGS__; break; }   \
        case 5: { constexpr int D = 5; __VA_ARGS__; break; }   \
        case 7: { constexpr int D = 7; __VA_ARGS__; break; }   \
        case 11: { constexpr int D = 11; __VA_ARGS__; break; } \
        case 17: { constexpr int D = 17; __VA_ARGS__; break; } \
        default: return LSP_ERR_PARAM;             \
    }

}  // namespace lsp
