// Experiment: a*b products as multiply-only IMAD.WIDE (2 cycles on the heavy pipe, vs 4 with a 64-bit accumulate)
// plus ALU-pipe carry adds; the m*p products stay fused.  See tools/mulbench.cu MODE 7.
#pragma once
#include "fr.cuh"
namespace lsp_split {
using namespace lsp;
// (lo,hi) += x*y + CC, sets CC  -- product formed separately
#define SPL_MADWC_CC(lo, hi, x, y) asm volatile("{.reg .u64 t; .reg .u32 tl, th;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {tl,th}, t;\n\taddc.cc.u32 %0, %0, tl;\n\taddc.cc.u32 %1, %1, th;}" : "+r"(lo), "+r"(hi) : "r"(x), "r"(y))
#define SPL_MADW_CC(lo, hi, x, y) asm volatile("{.reg .u64 t; .reg .u32 tl, th;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {tl,th}, t;\n\tadd.cc.u32 %0, %0, tl;\n\taddc.cc.u32 %1, %1, th;}" : "+r"(lo), "+r"(hi) : "r"(x), "r"(y))
#define SPL_MADWC3_CC(lo, hi, x, y, ilo, ihi) asm volatile("{.reg .u64 t; .reg .u32 tl, th;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {tl,th}, t;\n\taddc.cc.u32 %0, %4, tl;\n\taddc.cc.u32 %1, %5, th;}" : "=r"(lo), "=r"(hi) : "r"(x), "r"(y), "r"(ilo), "r"(ihi))

__device__ __forceinline__ void row(uint32_t* X, uint32_t* Y, const uint32_t* Z, const uint32_t* v, uint32_t s) {
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(X[0]) : "r"(Z[1]));
    SPL_MADWC3_CC(Y[0], Y[1], v[1], s, Z[2], Z[3]);
    SPL_MADWC3_CC(Y[2], Y[3], v[3], s, Z[4], Z[5]);
    SPL_MADWC3_CC(Y[4], Y[5], v[5], s, Z[6], Z[7]);
    asm volatile("madc.lo.cc.u32 %0, %2, %3, 0;\n\tmadc.hi.u32 %1, %2, %3, 0;" : "=r"(Y[6]), "=r"(Y[7]) : "r"(v[7]), "r"(s));
    SPL_MADW_CC(X[0], X[1], v[0], s);
    SPL_MADWC_CC(X[2], X[3], v[2], s);
    SPL_MADWC_CC(X[4], X[5], v[4], s);
    SPL_MADWC_CC(X[6], X[7], v[6], s);
    LSP_ADDC0(Y[7]);
}
__device__ __forceinline__ Fr fr_mul_split(const Fr& a, const Fr& b) {
    uint32_t E[8], O[8], E2[8], O2[8];
    mont_row<true, 0>(E, O, nullptr, a.l, b.l[0]);
    mont_reduce_step(E, O);
    row(O, E2, E, a.l, b.l[1]);   mont_reduce_step(O, E2);
    row(E2, O2, O, a.l, b.l[2]);  mont_reduce_step(E2, O2);
    row(O2, E, E2, a.l, b.l[3]);  mont_reduce_step(O2, E);
    row(E, O, O2, a.l, b.l[4]);   mont_reduce_step(E, O);
    row(O, E2, E, a.l, b.l[5]);   mont_reduce_step(O, E2);
    row(E2, O2, O, a.l, b.l[6]);  mont_reduce_step(E2, O2);
    row(O2, E, E2, a.l, b.l[7]);  mont_reduce_step(O2, E);
    Fr r = mont_finish(O2, E);
    fr_reduce_once(r);
    return r;
}
}  // namespace lsp_split
