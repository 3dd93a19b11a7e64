// Carry-free Montgomery product for BLS12-377 Fr on 9 x 29-bit limbs -- MEASURED ALTERNATIVE, not
// on the product path (tools/mulbench.cu, tools/permlat.cu): 45-47 G modmul/s against 57 G for the
// carry-chained 8 x 32-bit product of fr.cuh, and 1303 vs 997 cycles for a lone warp, because the
// ~180 shift/mask/carry instructions it adds cost as much issue time as the carries it removes.
//
// Every 29x29-bit product is < 2^58, so the 17 column accumulators of a 9x9 schoolbook
// product plus the 9x8 reduction products stay below 2^63: no carry flag is ever needed
// and every multiply is a plain IMAD.WIDE.U32 (the carry-chained IMAD.WIDE.U32.X of the
// 8 x 32-bit form issues at half rate on sm_100a -- tools/int_peak*.cu).
// The Montgomery radix stays R = 2^256 so values interoperate with `Fr`: eight reduction
// steps of 29 bits and one of 24 bits.  r = 1 (mod 2^47), so the quotient digit of each
// step is just the negated low bits of the running column.
#pragma once
#include "fr.cuh"

namespace lsp {

struct F29 {
    uint32_t l[9];  // limbs < 2^29; value = sum l[i] 2^(29 i)
};

constexpr uint32_t M29 = 0x1fffffffu;
#define LSP_P29_1 0x108c0000u
#define LSP_P29_2 0x00000042u
#define LSP_P29_3 0x14edfda0u
#define LSP_P29_4 0x1b00159au
#define LSP_P29_5 0x068f2e1bu
#define LSP_P29_6 0x155982d1u
#define LSP_P29_7 0x0bd34594u
#define LSP_P29_8 0x0012ab65u

__device__ __forceinline__ F29 f29_unpack(const Fr& a) {
    F29 r;
    r.l[0] = a.l[0] & M29;
    r.l[1] = __funnelshift_r(a.l[0], a.l[1], 29) & M29;
    r.l[2] = __funnelshift_r(a.l[1], a.l[2], 26) & M29;
    r.l[3] = __funnelshift_r(a.l[2], a.l[3], 23) & M29;
    r.l[4] = __funnelshift_r(a.l[3], a.l[4], 20) & M29;
    r.l[5] = __funnelshift_r(a.l[4], a.l[5], 17) & M29;
    r.l[6] = __funnelshift_r(a.l[5], a.l[6], 14) & M29;
    r.l[7] = __funnelshift_r(a.l[6], a.l[7], 11) & M29;
    r.l[8] = a.l[7] >> 8;
    return r;
}

// value < 2^256 required (true for every lazy result: < 2r)
__device__ __forceinline__ Fr f29_pack_lazy(const F29& a) {
    Fr r;
    r.l[0] = a.l[0] | (a.l[1] << 29);
    r.l[1] = (a.l[1] >> 3) | (a.l[2] << 26);
    r.l[2] = (a.l[2] >> 6) | (a.l[3] << 23);
    r.l[3] = (a.l[3] >> 9) | (a.l[4] << 20);
    r.l[4] = (a.l[4] >> 12) | (a.l[5] << 17);
    r.l[5] = (a.l[5] >> 15) | (a.l[6] << 14);
    r.l[6] = (a.l[6] >> 18) | (a.l[7] << 11);
    r.l[7] = (a.l[7] >> 21) | (a.l[8] << 8);
    return r;
}
__device__ __forceinline__ Fr f29_pack_canonical(const F29& a) {
    Fr r = f29_pack_lazy(a);
    fr_reduce_once(r);
    return r;
}

// Montgomery reduction of the 17 columns t[] (value sum t[k] 2^(29k) < 2^256 * 2r) to a
// normalised F29 < 2r.
__device__ __forceinline__ F29 f29_reduce(unsigned long long* t) {
    const uint32_t p1 = LSP_P29_1, p2 = LSP_P29_2, p3 = LSP_P29_3, p4 = LSP_P29_4, p5 = LSP_P29_5, p6 = LSP_P29_6,
                   p7 = LSP_P29_7, p8 = LSP_P29_8;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const uint32_t mask = (i < 8) ? M29 : 0x00ffffffu;
        const uint32_t m = (0u - uint32_t(t[i])) & mask;
        t[i + 1] += (unsigned long long)m * p1;
        t[i + 2] += (unsigned long long)m * p2;
        t[i + 3] += (unsigned long long)m * p3;
        t[i + 4] += (unsigned long long)m * p4;
        t[i + 5] += (unsigned long long)m * p5;
        t[i + 6] += (unsigned long long)m * p6;
        t[i + 7] += (unsigned long long)m * p7;
        t[i + 8] += (unsigned long long)m * p8;
        t[i] += m;  // low 29 (24) bits are now zero
        if (i < 8) t[i + 1] += t[i] >> 29;
    }
    // value = (sum_{k>=8} t[k] 2^(29(k-8))) >> 24
#pragma unroll
    for (int k = 8; k < 16; k++) {
        t[k + 1] += t[k] >> 29;
        t[k] &= M29;
    }
    F29 r;
#pragma unroll
    for (int j = 0; j < 8; j++) r.l[j] = ((uint32_t(t[8 + j]) >> 24) | (uint32_t(t[9 + j]) << 5)) & M29;
    r.l[8] = uint32_t(t[16] >> 24);
    return r;
}

// a, b normalised (limbs < 2^29), values < 2^256.  Result normalised, < 2r when a*b < 2^256 * r.
__device__ __forceinline__ F29 f29_mul(const F29& a, const F29& b) {
    unsigned long long t[17];
#pragma unroll
    for (int k = 0; k < 17; k++) t[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) t[i + j] += (unsigned long long)a.l[i] * b.l[j];
    return f29_reduce(t);
}

__device__ __forceinline__ F29 f29_sqr(const F29& a) {
    unsigned long long t[17];
#pragma unroll
    for (int k = 0; k < 17; k++) t[k] = 0;
    uint32_t d[9];
#pragma unroll
    for (int i = 0; i < 9; i++) d[i] = a.l[i] << 1;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        t[2 * i] += (unsigned long long)a.l[i] * a.l[i];
#pragma unroll
        for (int j = i + 1; j < 9; j++) t[i + j] += (unsigned long long)d[i] * a.l[j];
    }
    return f29_reduce(t);
}

}  // namespace lsp
