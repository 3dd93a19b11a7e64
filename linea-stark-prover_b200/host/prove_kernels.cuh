// Small kernels shared by the single-GPU prover (prover.cu) and the sharded one (sharded.cu).
// Anonymous namespace: each translation unit gets its own copy (the library is built without
// relocatable device code).
#pragma once
#include "../csrc/stark.cuh"

namespace {
using namespace lsp;

// Scalars the reduced-opening kernel needs, from the opened values:
//   s[0] = sum_i a^i y_zeta[i]   s[1] = sum_i a^i y_zeta'[i]   s[2] = sum_c a^c yq[c]
//   s[3] = a^W                   s[4] = a^(2W)
__global__ void k_open_scalars(const Fr* __restrict__ alpha, const Fr* __restrict__ y_zeta, const Fr* __restrict__ y_next,
                               const Fr* __restrict__ yq, int width, int q, Fr* __restrict__ s) {
    Fr a = fr_load(alpha);
    auto horner = [&](const Fr* y, int n) {
        Fr acc = fr_load(y + n - 1);
        for (int i = n - 2; i >= 0; i--) acc = fr_add(fr_mul(acc, a), fr_load(y + i));
        return acc;
    };
    fr_store(s + 0, horner(y_zeta, width));
    fr_store(s + 1, horner(y_next, width));
    fr_store(s + 2, horner(yq, q));
    Fr aw = fr_pow_u32(a, uint32_t(width));
    fr_store(s + 3, aw);
    fr_store(s + 4, fr_sqr(aw));
}

// final_poly = idft(bit_reverse(folded)); F <= 1024 values, one thread per coefficient.
__global__ void k_final_poly(const FieldConsts* __restrict__ fc, const Fr* __restrict__ folded, int log_f, Fr scale /* 1/F */, Fr* __restrict__ out) {
    int k = threadIdx.x;
    int f = 1 << log_f;
    if (k >= f) return;
    Fr w = fr_two_adic_generator(fc, log_f);
    Fr wk = fr_pow_u32(w, uint32_t((f - k) & (f - 1)));  // w^-k
    Fr acc = fr_zero(), wp = fr_one();
    for (int j = 0; j < f; j++) {
        acc = fr_add(acc, fr_mul(fr_load(folded + bitrev32(uint32_t(j), log_f)), wp));
        wp = fr_mul(wp, wk);
    }
    fr_store(out + k, fr_mul(acc, scale));
}

__global__ void k_make_cols(const Fr* base, size_t stride, int n, const Fr** out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) out[c] = base + size_t(c) * stride;
}

__global__ void k_set_small(Fr* dst, uint32_t v) {  // dst = Fr::from_canonical(v)
    Fr x = fr_zero();
    x.l[0] = v;
    fr_store(dst, fr_mul(x, fr_const(FR_R2)));
}

}  // namespace
