// First-generation Montgomery product (round 1a): kept for tools/mulbench.cu as the comparison
// point.  ptxas folds m = -X[0] into the constant products and then cannot fuse the m*p lanes
// into IMAD.WIDE.U32.X (48 IMAD.X + 64 IMAD.HI.U32 remain): 57 G modmul/s on B200.
#pragma once
#include "fr.cuh"
namespace lsp_v1 {
using lsp::Fr;
// ---- Montgomery product ---------------------------------------------------
// lane (lo,hi) = x*y                       (no carries)
// first lane of a chain: (lo,hi) += x*y, sets CC
// middle lane: (lo,hi) += x*y + CC, sets CC
// shifted variants: (lo,hi) = x*y + (ilo,ihi) [+ CC], sets CC

// One word-serial step on (X aligned at 2^0, Y aligned at 2^32).
//   FIRST: X = a_even*bi, Y = a_odd*bi.
//   else : the previous step left X_old[0] == 0; divide by 2^32 by renaming
//          (new X = old Y, new Y = old X >> 64, stray limb X_old[1] joins
//          new X[0]) while adding a*bi; then add m*p with m = -X[0].
// Caller passes the arrays already swapped: X = old Y, Z = old X.
template <bool FIRST>
__device__ __forceinline__ void mont_step_v1(uint32_t* X, uint32_t* Y, const uint32_t* Z, const uint32_t* a, uint32_t bi) {
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
__device__ __forceinline__ Fr fr_mul_lazy_v1(const Fr& a, const Fr& b) {
    uint32_t E[8], O[8];
    mont_step_v1<true>(E, O, nullptr, a.l, b.l[0]);
    // after step k the accumulator with X[0]==0 is the one passed as X
    uint32_t E2[8], O2[8];
    mont_step_v1<false>(O, E2, E, a.l, b.l[1]);   // X=O, new Y=E2 from old X=E
    mont_step_v1<false>(E2, O2, O, a.l, b.l[2]);
    mont_step_v1<false>(O2, E, E2, a.l, b.l[3]);
    mont_step_v1<false>(E, O, O2, a.l, b.l[4]);
    mont_step_v1<false>(O, E2, E, a.l, b.l[5]);
    mont_step_v1<false>(E2, O2, O, a.l, b.l[6]);
    mont_step_v1<false>(O2, E, E2, a.l, b.l[7]);
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


__device__ __forceinline__ Fr fr_mul_v1(const Fr& a, const Fr& b) {
    Fr r = fr_mul_lazy_v1(a, b);
    lsp::fr_reduce_once(r);
    return r;
}
}  // namespace lsp_v1
