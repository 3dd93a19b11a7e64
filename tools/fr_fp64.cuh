// BLS12-377 Fr on the FP64 pipe: 11 signed limbs of 24 bits held in doubles, Montgomery radix 2^264.
//
// MEASURED ALTERNATIVE, not on the product path (tools/fpmul.cu, profiles/r2d_fp64_pipe.log): exact and correct,
// but on B200 FP64 and integer-multiply work do NOT overlap -- see the result block at the end of this comment.
//
// Why: every kernel of this library is bound by the integer multiplier (IMAD.WIDE: 4 cycles per warp per SM
// sub-partition) while B200's FP64 pipe (DFMA: 2 cycles per warp, 64 per clock per SM) sits idle.  A DFMA on
// integers below 2^53 is an exact integer multiply-add, so a second big-integer multiplier is available for free
// as long as the work is expressed in pieces that fit 53 bits:
//   * limbs are balanced 24-bit digits (|x_k| <= 2^23 when normalised), products are < 2^46 and a column of a
//     word-serial Montgomery product accumulates 22 of them: < 2^51, exact;
//   * p = 1 (mod 2^24), so -p^{-1} = -1: the quotient digit of a step is q = -(t_0 mod+- 2^24), obtained with the
//     round-to-nearest trick (add and subtract 1.5 * 2^76), and t_0 + q*p_0 is just the rounded value;
//   * balanced digits make the Montgomery result a SIGNED residue with |r| <= p/2 + eps: no conditional
//     subtraction exists in this domain.
// Values are converted from / to the 8 x u32, R = 2^256 form at the edges of a permutation (one extra product
// each way to change the Montgomery radix).  This pipe does not exist in this form on B300 (FP64 is vestigial there).
//
// RESULT (B200, tools/fpmul.cu): bit-exact against fr_mul / the integer S-box on 16 k random pairs and on chained
// S-boxes.  Alone the FP64 S-box ((x+c)^5: 2 squares + 1 product + 1 normalisation, 875 DFMA/DADD) runs at
// 18.5 G S-box/s against 27.6 G for the integer one.  Run NEXT TO the integer kernel (two streams, 1-4 FP64 blocks
// + 4-6 integer blocks per SM) the total stays at 27-29 G S-box/s: the time the two kinds of work take simply
// adds up (e.g. 34 % of the cycles in DFMA + 70 % in IMAD.WIDE), although ncu counts them on different pipes and
// the issue slot is only 44 % used.  So the idle-looking FP64 pipe is not a second multiplier for this workload.
#pragma once
#include "fr.cuh"

namespace lsp {

constexpr int FP_N = 11;
struct Fp {
    double l[FP_N];
};

// balanced 24-bit digits of r
#define LSP_FP_P(k) ((k) == 0 ? 1.0 : (k) == 1 ? -8388608.0 : (k) == 2 ? 68114.0 : (k) == 3 ? -3145728.0 : (k) == 4 ? -5605633.0 : \
                     (k) == 5 ? -5242534.0 : (k) == 6 ? 1989688.0 : (k) == 7 ? 6337613.0 : (k) == 8 ? 2925910.0 : (k) == 9 ? 6643354.0 : 4779.0)

__device__ __forceinline__ double fp_rn24(double z) {  // nearest multiple of 2^24 (|z| < 2^75)
    const double M = 113336795588871485128704.0;       // 1.5 * 2^76
    return __dadd_rn(__dadd_rn(z, M), -M);
}
constexpr double FP_2M24 = 1.0 / 16777216.0;

// carry propagation to balanced digits: |l[k]| <= 2^23 for k < 10, the rest in l[10]
__device__ __forceinline__ void fp_normalize(Fp& x) {
#pragma unroll
    for (int k = 0; k < FP_N - 1; k++) {
        double h = fp_rn24(x.l[k]);
        x.l[k] = __dadd_rn(x.l[k], -h);
        x.l[k + 1] = __fma_rn(h, FP_2M24, x.l[k + 1]);
    }
}

__device__ __forceinline__ Fp fp_add(const Fp& a, const Fp& b) {  // limb-wise, no carries
    Fp r;
#pragma unroll
    for (int k = 0; k < FP_N; k++) r.l[k] = __dadd_rn(a.l[k], b.l[k]);
    return r;
}

// One reduction step on the window t[0..10]: clears t[0] modulo 2^24 with q*p and shifts the window down by a limb.
__device__ __forceinline__ void fp_reduce_row(double* t) {
    double h = fp_rn24(t[0]);
    double q = __dadd_rn(h, -t[0]);  // -(t0 mod+- 2^24) = t0 * (-1/p) mod+- 2^24
    t[0] = __fma_rn(h, FP_2M24, __fma_rn(q, LSP_FP_P(1), t[1]));
#pragma unroll
    for (int k = 2; k < FP_N; k++) t[k - 1] = __fma_rn(q, LSP_FP_P(k), t[k]);
    t[FP_N - 1] = 0.0;
}

// Montgomery product a*b/2^264, signed result with |r| <= p/2 + |a||b|/2^264; normalised digits out.
// Exactness needs |a.l[i] * b.l[k]| summed over a column (<= 11 terms, plus 11 terms q*p_k < 2^46) below 2^53:
// fine for |a.l| <= 2^25, |b.l| <= 2^24.
__device__ __forceinline__ Fp fp_mul(const Fp& a, const Fp& b) {
    double t[FP_N];
#pragma unroll
    for (int k = 0; k < FP_N; k++) t[k] = 0.0;
#pragma unroll
    for (int i = 0; i < FP_N; i++) {
#pragma unroll
        for (int k = 0; k < FP_N; k++) t[k] = __fma_rn(a.l[i], b.l[k], t[k]);
        fp_reduce_row(t);
    }
    Fp r;
#pragma unroll
    for (int k = 0; k < FP_N; k++) r.l[k] = t[k];
    fp_normalize(r);
    return r;
}

// Montgomery square: the 66 distinct products first (cross terms through the doubled operand), then 11 reduction
// rows over the 21-limb product.  Needs normalised input (|a.l| <= 2^23 + small).
__device__ __forceinline__ Fp fp_sqr(const Fp& a) {
    double t[2 * FP_N];
#pragma unroll
    for (int k = 0; k < 2 * FP_N; k++) t[k] = 0.0;
#pragma unroll
    for (int i = 0; i < FP_N; i++) {
        t[2 * i] = __fma_rn(a.l[i], a.l[i], t[2 * i]);
        const double a2 = __dadd_rn(a.l[i], a.l[i]);
#pragma unroll
        for (int j = i + 1; j < FP_N; j++) t[i + j] = __fma_rn(a2, a.l[j], t[i + j]);
    }
#pragma unroll
    for (int i = 0; i < FP_N; i++) {   // row i works on the window t[i .. i+10]; the shift is an index offset here
        double* w = t + i;
        double h = fp_rn24(w[0]);
        double q = __dadd_rn(h, -w[0]);
        w[1] = __fma_rn(h, FP_2M24, __fma_rn(q, LSP_FP_P(1), w[1]));
#pragma unroll
        for (int k = 2; k < FP_N; k++) w[k] = __fma_rn(q, LSP_FP_P(k), w[k]);
    }
    Fp r;
#pragma unroll
    for (int k = 0; k < FP_N; k++) r.l[k] = t[FP_N + k];
    fp_normalize(r);
    return r;
}

// ---- conversions -----------------------------------------------------------------------------------------
// 8 x u32 (any value < 2^256) -> 11 unsigned 24-bit digits as doubles (no change of Montgomery radix)
__device__ __forceinline__ Fp fp_from_limbs(const Fr& a) {
    Fp r;
#pragma unroll
    for (int k = 0; k < FP_N; k++) {
        const int bit = 24 * k, w = bit >> 5, sh = bit & 31;
        uint32_t lo = a.l[w], hi = (w + 1 < 8) ? a.l[w + 1] : 0u;
        uint32_t f = sh ? __funnelshift_r(lo, hi, sh) : lo;
        if (k < FP_N - 1) f &= 0xffffffu;   // the top digit holds bits 240..255
        r.l[k] = __uint2double_rn(f);
    }
    return r;
}
// signed value with |x| < p (normalised or not) -> canonical 8 x u32 in [0, p)
__device__ __forceinline__ Fr fp_to_limbs(Fp x) {
#pragma unroll
    for (int k = 0; k < FP_N; k++) x.l[k] = __dadd_rn(x.l[k], LSP_FP_P(k));   // + p: now in (0, 2p)
    fp_normalize(x);
    long long c = 0;
    uint32_t d[FP_N];
#pragma unroll
    for (int k = 0; k < FP_N; k++) {
        c += (long long)__double2ll_rn(x.l[k]);
        d[k] = k < FP_N - 1 ? (uint32_t)(c & 0xffffff) : (uint32_t)c;
        c >>= 24;
    }
    Fr r;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        // word w = bits [32w, 32w+32): digits k0 = floor(32w/24) ...
        const int bit = 32 * w, k0 = bit / 24, sh = bit - 24 * k0;
        unsigned long long v = (unsigned long long)d[k0] >> sh;
        v |= (unsigned long long)d[k0 + 1] << (24 - sh);
        if (k0 + 2 < FP_N) v |= (unsigned long long)d[k0 + 2] << (48 - sh);
        r.l[w] = (uint32_t)v;
    }
    fr_reduce_once(r);
    return r;
}

// Montgomery-radix change constants (balanced digits): K_IN = 2^8 * 2^264 mod p, K_OUT = 2^256 mod p
#define LSP_FP_KIN {-898642.0, 0.0, -6367899.0, 6287808.0, -2348415.0, 1845772.0, -1904905.0, -1253112.0, -3884947.0, 7459451.0, 1338.0}
#define LSP_FP_KOUT {-14.0, 0.0, -953589.0, -6291456.0, -5407215.0, 6286617.0, 5698804.0, -4840504.0, -7408313.0, 7656338.0, -1376.0}

// x*2^256 (canonical 8 x u32) -> x*2^264 (fp)         and back
__device__ __forceinline__ Fp fp_from_mont(const Fr& a) {
    const Fp kin = {LSP_FP_KIN};
    return fp_mul(fp_from_limbs(a), kin);
}
__device__ __forceinline__ Fr fp_to_mont(const Fp& x) {
    const Fp kout = {LSP_FP_KOUT};
    return fp_to_limbs(fp_mul(x, kout));
}

}  // namespace lsp
