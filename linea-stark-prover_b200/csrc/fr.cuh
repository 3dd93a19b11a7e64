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
// Cost model measured on B200 (profiles/r1*_k_leaf_hash, tools/pipe_probe.cu): every 32x32->64
// product occupies the FMA-heavy pipe for 4 cycles per warp as one IMAD.WIDE.U32(.X), or 6 as an
// IMAD (lo) + IMAD.HI pair; IADD3/SHF/SEL go to the ALU pipe.  So the kernels here are bound by
// the NUMBER OF IMAD.WIDE: a product is 64 (a*b) + 56 (m*p: p[0] = 1 needs no multiply) = 120,
// a square 36 + 56 = 92.
// lane (lo,hi) = x*y                       (no carries)
#define LSP_MULW(lo, hi, x, y) asm("{.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0,%1}, t;}" : "=r"(lo), "=r"(hi) : "r"(x), "r"(y))
// first lane of a chain: (lo,hi) += x*y, sets CC
#define LSP_MADW_CC(lo, hi, x, y) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(y))
// middle lane: (lo,hi) += x*y + CC, sets CC
#define LSP_MADWC_CC(lo, hi, x, y) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(y))
// shifted variants: (lo,hi) = x*y + (ilo,ihi) [+ CC], sets CC
#define LSP_MADW3_CC(lo, hi, x, y, ilo, ihi) asm volatile("mad.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(x), "r"(y), "r"(ilo), "r"(ihi))
#define LSP_MADWC3_CC(lo, hi, x, y, ilo, ihi) asm volatile("madc.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(x), "r"(y), "r"(ilo), "r"(ihi))
// (lo,hi) = (ilo,ihi) + CC, sets CC   (a lane that receives no product this step)
#define LSP_COPYC_CC(lo, hi, ilo, ihi) asm volatile("addc.cc.u32 %0, %2, 0;\n\taddc.cc.u32 %1, %3, 0;" : "=r"(lo), "=r"(hi) : "r"(ilo), "r"(ihi))
#define LSP_ADDC0(x) asm volatile("addc.u32 %0, %0, 0;" : "+r"(x))

// The two accumulators of the word-serial product: X holds the limbs aligned at 2^0, Y those
// aligned at 2^32, so every 32x32->64 product lands on a 64-bit lane (an even/odd register pair)
// and ptxas emits one IMAD.WIDE.U32(.X) per product with the carry in a predicate.
//
// Reduction half of a step: m = -X[0] (-r^{-1} mod 2^32 = 0xffffffff for this modulus), then
// X + 2^32 Y += m * p, which clears X[0].  m comes out of an asm statement on purpose: when ptxas
// sees m = -X[0] it rewrites m*p_i as X[0]*(-p_i) and then splits every one of these lanes into
// IMAD + IMAD.HI.U32 (6 pipe cycles instead of 4).  p[0] = 1: that lane is an addition.
// Pipe steering: ptxas implements m = -X[0] as IMAD.MOV (FMA-heavy pipe, 2 cycles there); the
// form below (an add whose carry-out is consumed, then a NOT) stays on the ALU pipe.  The X and
// Y carry chains are kept independent on purpose: writing the product as ONE unbroken chain
// keeps every carry add on the ALU pipe too, but serialises ~180 instructions and was slower
// (leaf hash 36.3 -> 38.2 ms, lone-warp permutation 43 -> 56 us).
__device__ __forceinline__ void mont_reduce_step(uint32_t* X, uint32_t* Y) {
    uint32_t m, d;
    const uint32_t p1 = LSP_P1, p2 = LSP_P2, p3 = LSP_P3, p4 = LSP_P4, p5 = LSP_P5, p6 = LSP_P6, p7 = LSP_P7;
    // d = X[0] - 1 carries exactly when X[0] != 0, which is the carry of X[0] + m*p[0] into X[1];
    // m = -X[0] = ~(X[0] - 1).  (CC.CF is the adder's carry, also after sub.cc: not a borrow flag.)
    asm volatile("add.cc.u32 %0, %2, 0xffffffff;\n\taddc.cc.u32 %1, %1, 0;" : "=r"(d), "+r"(X[1]) : "r"(X[0]));
    asm volatile("not.b32 %0, %1;" : "=r"(m) : "r"(d));
    LSP_MADWC_CC(X[2], X[3], p2, m);
    LSP_MADWC_CC(X[4], X[5], p4, m);
    LSP_MADWC_CC(X[6], X[7], p6, m);
    LSP_ADDC0(Y[7]);
    LSP_MADW_CC(Y[0], Y[1], p1, m);
    LSP_MADWC_CC(Y[2], Y[3], p3, m);
    LSP_MADWC_CC(Y[4], Y[5], p5, m);
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(Y[6]), "+r"(Y[7]) : "r"(p7), "r"(m));
}

// Product half of a step on (X aligned at 2^0, Y aligned at 2^32): adds v * s where lane j of v
// has weight 2^(32 j); lanes below FROM are absent (squaring rows).
//   FIRST: X = v_even*s, Y = v_odd*s.
//   else : the previous step left X_old[0] == 0; divide by 2^32 by renaming (new X = old Y,
//          new Y = old X >> 64, stray limb X_old[1] joins new X[0]) while adding v*s.
// Caller passes the arrays already swapped: X = old Y, Z = old X.
template <bool FIRST, int FROM>
__device__ __forceinline__ void mont_row(uint32_t* X, uint32_t* Y, const uint32_t* Z, const uint32_t* v, uint32_t s) {
    if (FIRST) {
        LSP_MULW(X[0], X[1], v[0], s);
        LSP_MULW(X[2], X[3], v[2], s);
        LSP_MULW(X[4], X[5], v[4], s);
        LSP_MULW(X[6], X[7], v[6], s);
        LSP_MULW(Y[0], Y[1], v[1], s);
        LSP_MULW(Y[2], Y[3], v[3], s);
        LSP_MULW(Y[4], Y[5], v[5], s);
        LSP_MULW(Y[6], Y[7], v[7], s);
    } else {
        // stray limb, carry feeds the Y chain (weight 2^32)
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(X[0]) : "r"(Z[1]));
        if (1 >= FROM) LSP_MADWC3_CC(Y[0], Y[1], v[1], s, Z[2], Z[3]); else LSP_COPYC_CC(Y[0], Y[1], Z[2], Z[3]);
        if (3 >= FROM) LSP_MADWC3_CC(Y[2], Y[3], v[3], s, Z[4], Z[5]); else LSP_COPYC_CC(Y[2], Y[3], Z[4], Z[5]);
        if (5 >= FROM) LSP_MADWC3_CC(Y[4], Y[5], v[5], s, Z[6], Z[7]); else LSP_COPYC_CC(Y[4], Y[5], Z[6], Z[7]);
        asm volatile("madc.lo.cc.u32 %0, %2, %3, 0;\n\tmadc.hi.u32 %1, %2, %3, 0;" : "=r"(Y[6]), "=r"(Y[7]) : "r"(v[7]), "r"(s));
        if (0 >= FROM) LSP_MADW_CC(X[0], X[1], v[0], s);
        if (2 >= FROM) { if (0 < FROM) LSP_MADW_CC(X[2], X[3], v[2], s); else LSP_MADWC_CC(X[2], X[3], v[2], s); }
        if (4 >= FROM) { if (2 < FROM) LSP_MADW_CC(X[4], X[5], v[4], s); else LSP_MADWC_CC(X[4], X[5], v[4], s); }
        if (6 >= FROM) { if (4 < FROM) LSP_MADW_CC(X[6], X[7], v[6], s); else LSP_MADWC_CC(X[6], X[7], v[6], s); }
        if (6 >= FROM) LSP_ADDC0(Y[7]);
    }
}

// T = X + Y*2^32 with X[0] == 0; result = T / 2^32.
__device__ __forceinline__ Fr mont_finish(const uint32_t* X, const uint32_t* Y) {
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
        : "r"(Y[0]), "r"(Y[1]), "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]),
          "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]));
    return r;
}

// Montgomery product, result in [0, 2r) provided a < 2^255 (b arbitrary < 2^256)
// and a*b < 2^256 * r  (true for a, b < 3r).
__device__ __forceinline__ Fr fr_mul_lazy(const Fr& a, const Fr& b) {
    uint32_t E[8], O[8], E2[8], O2[8];
    mont_row<true, 0>(E, O, nullptr, a.l, b.l[0]);
    mont_reduce_step(E, O);
    // after step k the accumulator with X[0]==0 is the one passed as X
    mont_row<false, 0>(O, E2, E, a.l, b.l[1]);   // X=O, new Y=E2 from old X=E
    mont_reduce_step(O, E2);
    mont_row<false, 0>(E2, O2, O, a.l, b.l[2]);
    mont_reduce_step(E2, O2);
    mont_row<false, 0>(O2, E, E2, a.l, b.l[3]);
    mont_reduce_step(O2, E);
    mont_row<false, 0>(E, O, O2, a.l, b.l[4]);
    mont_reduce_step(E, O);
    mont_row<false, 0>(O, E2, E, a.l, b.l[5]);
    mont_reduce_step(O, E2);
    mont_row<false, 0>(E2, O2, O, a.l, b.l[6]);
    mont_reduce_step(E2, O2);
    mont_row<false, 0>(O2, E, E2, a.l, b.l[7]);
    mont_reduce_step(O2, E);
    return mont_finish(O2, E);
}

// Montgomery square, result in [0, 2r) for a < 2^253.  Row k adds a_k * (a_k 2^(32k) +
// 2 * sum_{j>k} a_j 2^(32j)): the doubled tail is read from the limbs of 2a (which fits 8 limbs),
// except that its lowest limb drops the bit shifted in from a_k.  36 products instead of 64; every
// column is complete by the time its reduction step runs because row k only touches columns >= 2k.
template <int K>
__device__ __forceinline__ void sqr_row(uint32_t* X, uint32_t* Y, const uint32_t* Z, const Fr& a, const uint32_t* sh, const uint32_t* db) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = j == K ? a.l[j] : (j == K + 1 ? sh[j] : db[j]);
    mont_row<K == 0, K>(X, Y, Z, v, a.l[K]);
}
__device__ __forceinline__ Fr fr_sqr_lazy(const Fr& a) {
    uint32_t sh[8], db[8];
    sh[0] = db[0] = 0;
#pragma unroll
    for (int j = 1; j < 8; j++) {
        asm("shf.l.clamp.b32 %0, 0, %1, 1;" : "=r"(sh[j]) : "r"(a.l[j]));  // a << 1 on the ALU pipe (ptxas would pick IMAD.IADD a+a)
        db[j] = __funnelshift_l(a.l[j - 1], a.l[j], 1);
    }
    uint32_t E[8], O[8], E2[8], O2[8];
    sqr_row<0>(E, O, nullptr, a, sh, db);
    mont_reduce_step(E, O);
    sqr_row<1>(O, E2, E, a, sh, db);
    mont_reduce_step(O, E2);
    sqr_row<2>(E2, O2, O, a, sh, db);
    mont_reduce_step(E2, O2);
    sqr_row<3>(O2, E, E2, a, sh, db);
    mont_reduce_step(O2, E);
    sqr_row<4>(E, O, O2, a, sh, db);
    mont_reduce_step(E, O);
    sqr_row<5>(O, E2, E, a, sh, db);
    mont_reduce_step(O, E2);
    sqr_row<6>(E2, O2, O, a, sh, db);
    mont_reduce_step(E2, O2);
    sqr_row<7>(O2, E, E2, a, sh, db);
    mont_reduce_step(O2, E);
    return mont_finish(O2, E);
}

__device__ __forceinline__ Fr fr_mul(const Fr& a, const Fr& b) {
    Fr r = fr_mul_lazy(a, b);
    fr_reduce_once(r);
    return r;
}
__device__ __forceinline__ Fr fr_sqr(const Fr& a) {
    Fr r = fr_sqr_lazy(a);
    fr_reduce_once(r);
    return r;
}

// Out-of-line product (arguments and result travel in registers, ~20 MOVs per call).
// Used where many products follow each other (Poseidon2 S-boxes, exponentiations): one
// 4 KiB body shared by every call site keeps those loops inside the instruction cache --
// with the product inlined a Poseidon2 permutation is ~84 KiB of straight-line code and
// the leaf-hash kernel stalls on instruction fetch (profiles/r1a: stall_no_instruction).
static __device__ __noinline__ Fr fr_mul_call(Fr a, Fr b) { return fr_mul(a, b); }
static __device__ __noinline__ Fr fr_sqr_call(Fr a) { return fr_sqr(a); }

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
