// BLS12-377 scalar field Fr for sm_100a -- 8 x u32 limbs, Montgomery form, R = 2^256.
//
// Replaces (on the device) the arkworks arithmetic behind the reference's
// `Val = Bls12_377Fr` (reference bin/src/config.rs:9; ark-ff 0.5.0
// `Fp256<MontBackend<FrConfig,4>>`, Cargo.lock:83-85).  The in-memory format
// is identical to the host type: 4 x u64 little-endian limbs of a*R mod r,
// fully reduced, so matrices cross the FFI boundary without conversion.
//
// Multiplication is a word-serial Montgomery product on two interleaved
// accumulators ("even" limbs aligned at 2^0, "odd" limbs aligned at 2^32) so
// that every 32x32->64 product lands on a 64-bit lane boundary and ptxas can
// emit one IMAD.WIDE.U32(.X) per product with the carry in a predicate.
// -r^{-1} mod 2^32 = 0xffffffff for this modulus, so the quotient digit is a
// plain negation.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lsp {

struct __align__(16) Fr {
    uint32_t l[8];
};

// modulus r, little-endian u32 limbs
#define LSP_P0 0x00000001u
#define LSP_P1 0x0a118000u
#define LSP_P2 0xd0000001u
#define LSP_P3 0x59aa76feu
#define LSP_P4 0x5c37b001u
#define LSP_P5 0x60b44d1eu
#define LSP_P6 0x9a2ca556u
#define LSP_P7 0x12ab655eu

__device__ __constant__ const uint32_t FR_P[8] = {LSP_P0, LSP_P1, LSP_P2, LSP_P3, LSP_P4, LSP_P5, LSP_P6, LSP_P7};
// R mod r (Montgomery one) and R^2 mod r
__device__ __constant__ const uint32_t FR_ONE[8] = {0xfffffff3u, 0x7d1c7fffu, 0x6ffffff2u, 0x7257f50fu,
                                                    0x512c0feeu, 0x16d81575u, 0x2bbb9a9du, 0x0d4bda32u};
__device__ __constant__ const uint32_t FR_R2[8] = {0xb861857bu, 0x25d577bau, 0x8860591fu, 0xcc2c27b5u,
                                                   0xe5dc8593u, 0xa7cc008fu, 0xeff1c939u, 0x011fdae7u};

__device__ __forceinline__ Fr fr_zero() {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
    return r;
}
__device__ __forceinline__ Fr fr_one() {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = FR_ONE[i];
    return r;
}
__device__ __forceinline__ Fr fr_load(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ Fr fr_load_nc(const Fr* p) {  // read-only path
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void fr_store(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ bool fr_is_zero(const Fr& a) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.l[i];
    return t == 0;
}
__device__ __forceinline__ bool fr_eq(const Fr& a, const Fr& b) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.l[i] ^ b.l[i];
    return t == 0;
}

// ---- 256-bit helpers ------------------------------------------------------
// r = a + b, returns nothing (inputs < 2^255 so no carry-out is possible)
__device__ __forceinline__ void u256_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    asm volatile("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}
// r = a - b, returns borrow (1 if a < b)
__device__ __forceinline__ uint32_t u256_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t bw;
    asm volatile("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return bw;  // 0 or 0xffffffff
}
// a in [0, 2r) -> [0, r)
__device__ __forceinline__ void fr_reduce_once(Fr& a) {
    const uint32_t p[8] = {LSP_P0, LSP_P1, LSP_P2, LSP_P3, LSP_P4, LSP_P5, LSP_P6, LSP_P7};
    uint32_t t[8];
    uint32_t bw = u256_sub(t, a.l, p);
#pragma unroll
    for (int i = 0; i < 8; i++) a.l[i] = bw ? a.l[i] : t[i];
}
__device__ __forceinline__ Fr fr_add(const Fr& a, const Fr& b) {
    Fr r;
    u256_add(r.l, a.l, b.l);
    fr_reduce_once(r);
    return r;
}
__device__ __forceinline__ Fr fr_sub(const Fr& a, const Fr& b) {
    const uint32_t p[8] = {LSP_P0, LSP_P1, LSP_P2, LSP_P3, LSP_P4, LSP_P5, LSP_P6, LSP_P7};
    Fr r;
    uint32_t bw = u256_sub(r.l, a.l, b.l);
    uint32_t t[8];
    u256_add(t, r.l, p);
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = bw ? t[i] : r.l[i];
    return r;
}
__device__ __forceinline__ Fr fr_neg(const Fr& a) {
    const uint32_t p[8] = {LSP_P0, LSP_P1, LSP_P2, LSP_P3, LSP_P4, LSP_P5, LSP_P6, LSP_P7};
    Fr r;
    u256_sub(r.l, p, a.l);
    bool z = fr_is_zero(a);
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = z ? 0u : r.l[i];
    return r;
}
__device__ __forceinline__ Fr fr_dbl(const Fr& a) { return fr_add(a, a); }
// a/2 mod r: (a + (a odd ? r : 0)) >> 1
__device__ __forceinline__ Fr fr_halve(const Fr& a) {
    const uint32_t p[8] = {LSP_P0, LSP_P1, LSP_P2, LSP_P3, LSP_P4, LSP_P5, LSP_P6, LSP_P7};
    uint32_t m = 0u - (a.l[0] & 1u);
    uint32_t q[8], t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = p[i] & m;
    u256_add(t, a.l, q);  // < 2^254, no carry-out
    Fr r;
#pragma unroll
    for (int i = 0; i < 7; i++) r.l[i] = (t[i] >> 1) | (t[i + 1] << 31);
    r.l[7] = t[7] >> 1;
    return r;
}

// ---- Montgomery product ---------------------------------------------------
// lane (lo,hi) = x*y                       (no carries)
#define LSP_MULW(lo, hi, x, y) asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(x), "r"(y))
// first lane of a chain: (lo,hi) += x*y, sets CC
#define LSP_MADW_CC(lo, hi, x, y) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(y))
// middle lane: (lo,hi) += x*y + CC, sets CC
#define LSP_MADWC_CC(lo, hi, x, y) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(y))
// shifted variants: (lo,hi) = x*y + (ilo,ihi) [+ CC], sets CC
#define LSP_MADW3_CC(lo, hi, x, y, ilo, ihi) asm volatile("mad.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(x), "r"(y), "r"(ilo), "r"(ihi))
#define LSP_MADWC3_CC(lo, hi, x, y, ilo, ihi) asm volatile("madc.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(x), "r"(y), "r"(ilo), "r"(ihi))
#define LSP_ADDC0(x) asm volatile("addc.u32 %0, %0, 0;" : "+r"(x))

// One word-serial step on (X aligned at 2^0, Y aligned at 2^32).
//   FIRST: X = a_even*bi, Y = a_odd*bi.
//   else : the previous step left X_old[0] == 0; divide by 2^32 by renaming
//          (new X = old Y, new Y = old X >> 64, stray limb X_old[1] joins
//          new X[0]) while adding a*bi; then add m*p with m = -X[0].
// Caller passes the arrays already swapped: X = old Y, Z = old X.
template <bool FIRST>
__device__ __forceinline__ void mont_step(uint32_t* X, uint32_t* Y, const uint32_t* Z, const uint32_t* a, uint32_t bi) {
    if (FIRST) {
        LSP_MULW(X[0], X[1], a[0], bi);
        LSP_MULW(X[2], X[3], a[2], bi);
        LSP_MULW(X[4], X[5], a[4], bi);
        LSP_MULW(X[6], X[7], a[6], bi);
        LSP_MULW(Y[0], Y[1], a[1], bi);
        LSP_MULW(Y[2], Y[3], a[3], bi);
        LSP_MULW(Y[4], Y[5], a[5], bi);
        LSP_MULW(Y[6], Y[7], a[7], bi);
    } else {
        // stray limb, carry feeds the Y chain (weight 2^32)
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(X[0]) : "r"(Z[1]));
        LSP_MADWC3_CC(Y[0], Y[1], a[1], bi, Z[2], Z[3]);
        LSP_MADWC3_CC(Y[2], Y[3], a[3], bi, Z[4], Z[5]);
        LSP_MADWC3_CC(Y[4], Y[5], a[5], bi, Z[6], Z[7]);
        asm volatile("madc.lo.cc.u32 %0, %2, %3, 0;\n\tmadc.hi.u32 %1, %2, %3, 0;" : "=r"(Y[6]), "=r"(Y[7]) : "r"(a[7]), "r"(bi));
        LSP_MADW_CC(X[0], X[1], a[0], bi);
        LSP_MADWC_CC(X[2], X[3], a[2], bi);
        LSP_MADWC_CC(X[4], X[5], a[4], bi);
        LSP_MADWC_CC(X[6], X[7], a[6], bi);
        LSP_ADDC0(Y[7]);
    }
    uint32_t m = 0u - X[0];
    const uint32_t p0 = LSP_P0, p1 = LSP_P1, p2 = LSP_P2, p3 = LSP_P3, p4 = LSP_P4, p5 = LSP_P5, p6 = LSP_P6, p7 = LSP_P7;
    LSP_MADW_CC(Y[0], Y[1], p1, m);
    LSP_MADWC_CC(Y[2], Y[3], p3, m);
    LSP_MADWC_CC(Y[4], Y[5], p5, m);
    LSP_MADWC_CC(Y[6], Y[7], p7, m);
    LSP_MADW_CC(X[0], X[1], p0, m);
    LSP_MADWC_CC(X[2], X[3], p2, m);
    LSP_MADWC_CC(X[4], X[5], p4, m);
    LSP_MADWC_CC(X[6], X[7], p6, m);
    LSP_ADDC0(Y[7]);
}

// Montgomery product, result in [0, 2r) provided a < 2^255 (b arbitrary < 2^256)
// and a*b < 2^256 * r  (true for a, b < 3r).
__device__ __forceinline__ Fr fr_mul_lazy(const Fr& a, const Fr& b) {
    uint32_t E[8], O[8];
    mont_step<true>(E, O, nullptr, a.l, b.l[0]);
    // after step k the accumulator with X[0]==0 is the one passed as X
    uint32_t E2[8], O2[8];
    mont_step<false>(O, E2, E, a.l, b.l[1]);   // X=O, new Y=E2 from old X=E
    mont_step<false>(E2, O2, O, a.l, b.l[2]);
    mont_step<false>(O2, E, E2, a.l, b.l[3]);
    mont_step<false>(E, O, O2, a.l, b.l[4]);
    mont_step<false>(O, E2, E, a.l, b.l[5]);
    mont_step<false>(E2, O2, O, a.l, b.l[6]);
    mont_step<false>(O2, E, E2, a.l, b.l[7]);
    // T = X + Y*2^32 with X = O2 (X[0] == 0), Y = E.  result = T / 2^32.
    Fr r;
    asm volatile("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, 0;"
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
          "r"(O2[1]), "r"(O2[2]), "r"(O2[3]), "r"(O2[4]), "r"(O2[5]), "r"(O2[6]), "r"(O2[7]));
    return r;
}

__device__ __forceinline__ Fr fr_mul(const Fr& a, const Fr& b) {
    Fr r = fr_mul_lazy(a, b);
    fr_reduce_once(r);
    return r;
}
__device__ __forceinline__ Fr fr_sqr(const Fr& a) { return fr_mul(a, a); }

// Out-of-line product (arguments and result travel in registers, ~20 MOVs per call).
// Used where many products follow each other (Poseidon2 S-boxes, exponentiations): one
// 4 KiB body shared by every call site keeps those loops inside the instruction cache --
// with the product inlined a Poseidon2 permutation is ~84 KiB of straight-line code and
// the leaf-hash kernel stalls on instruction fetch (profiles/r1a: stall_no_instruction).
static __device__ __noinline__ Fr fr_mul_call(Fr a, Fr b) { return fr_mul(a, b); }
__device__ __forceinline__ Fr fr_sqr_call(const Fr& a) { return fr_mul_call(a, a); }

// a^e for a small runtime exponent (e < 2^32)
__device__ __forceinline__ Fr fr_pow_u32(Fr a, uint32_t e) {
    Fr r = fr_one();
    while (e) {
        if (e & 1u) r = fr_mul_call(r, a);
        a = fr_sqr_call(a);
        e >>= 1;
    }
    return r;
}

// a^(r-2) (Fermat); inverse of zero is zero.
static __device__ __noinline__ Fr fr_inv(const Fr& a) {
    // r - 2, little-endian u32 limbs (r ends in ...00000001)
    const uint32_t ex[8] = {0xffffffffu, LSP_P1 - 1u, LSP_P2, LSP_P3, LSP_P4, LSP_P5, LSP_P6, LSP_P7};
    Fr r = a;
    // top set bit of r-2 is bit 252 = bit 28 of limb 7; start below it
    for (int i = 7; i >= 0; i--) {
        for (int b = (i == 7 ? 27 : 31); b >= 0; b--) {
            r = fr_sqr_call(r);
            if ((ex[i] >> b) & 1u) r = fr_mul_call(r, a);
        }
    }
    return r;
}

}  // namespace lsp
