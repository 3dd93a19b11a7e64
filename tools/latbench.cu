// Lone-warp latencies (cycles, clock64) of the pieces of the three-lanes-per-permutation shape:
// product, square, S-box, lane-triple sum, canonicalisations, and the whole permutation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I linea-stark-prover_b200/csrc tools/latbench.cu -o tools/latbench
#include <cstdio>
#include <cstring>
#include "poseidon2.cuh"
using namespace lsp;


// Timing probe only (the result is NOT a Montgomery product): the reduction digit of each step comes from a value
// that has been ready for a whole step, so the X[0] -> m -> seven products dependency of the real code is gone
// while the instruction mix stays the same.
__device__ __forceinline__ void mont_reduce_step_fake(uint32_t* X, uint32_t* Y, uint32_t m_early) {
    uint32_t m, d;
    const uint32_t p1 = LSP_P1, p2 = LSP_P2, p3 = LSP_P3, p4 = LSP_P4, p5 = LSP_P5, p6 = LSP_P6, p7 = LSP_P7;
    asm volatile("add.cc.u32 %0, %2, 0xffffffff;\n\taddc.cc.u32 %1, %1, 0;" : "=r"(d), "+r"(X[1]) : "r"(m_early));
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
__device__ __forceinline__ Fr fr_mul_fake_m(const Fr& a, const Fr& b) {
    uint32_t E[8], O[8], E2[8], O2[8];
    mont_row<true, 0>(E, O, nullptr, a.l, b.l[0]);
    mont_reduce_step_fake(E, O, b.l[0]);
    mont_row<false, 0>(O, E2, E, a.l, b.l[1]);
    mont_reduce_step_fake(O, E2, E[2]);
    mont_row<false, 0>(E2, O2, O, a.l, b.l[2]);
    mont_reduce_step_fake(E2, O2, O[2]);
    mont_row<false, 0>(O2, E, E2, a.l, b.l[3]);
    mont_reduce_step_fake(O2, E, E2[2]);
    mont_row<false, 0>(E, O, O2, a.l, b.l[4]);
    mont_reduce_step_fake(E, O, O2[2]);
    mont_row<false, 0>(O, E2, E, a.l, b.l[5]);
    mont_reduce_step_fake(O, E2, E[2]);
    mont_row<false, 0>(E2, O2, O, a.l, b.l[6]);
    mont_reduce_step_fake(E2, O2, O[2]);
    mont_row<false, 0>(O2, E, E2, a.l, b.l[7]);
    mont_reduce_step_fake(O2, E, E2[2]);
    return mont_finish(O2, E);
}

template <int MODE>
__global__ void __launch_bounds__(32) k(const __grid_constant__ P2Params P, Fr* io, long long* clk, int reps) {
    const int lane = threadIdx.x, kk = lane / 3, w = lane - 3 * kk;
    Fr a = fr_load(io + w), b = fr_load(io + 3);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < reps; i++) {
        if (MODE == 0) a = fr_mul_lazy(a, b);
        if (MODE == 1) a = fr_sqr_lazy(a);
        if (MODE == 2) a = p2_sbox<5>(a);
        if (MODE == 3) a = fr_add_lazy(b, p2_tri_sum(a, 3 * kk));
        if (MODE == 4) { a = fr_add_lazy(a, b); fr_canon8(a); }
        if (MODE == 5) { a = fr_add_lazy(a, b); fr_reduce_once(a); }
        if (MODE == 6) p2_permute_tri<5>(P, a, w, 3 * kk);
        if (MODE == 7) a = fr_mul(a, b);
        if (MODE == 8) a = fr_mul_fake_m(a, b);
    }
    long long t1 = clock64();
    fr_store(io + 8 + lane, a);
    if (lane == 0) *clk = (t1 - t0) / reps;
}

int main() {
    P2Params P; memset(&P, 0, sizeof P);
    P.half_f = 4; P.rounds_p = 22; P.sbox_d = 5; P.diag_kind = 1;
    for (int r = 0; r < 4; r++) for (int i = 0; i < 3; i++) { P.ext_initial[r][i].l[0] = r * 3 + i + 1; P.ext_terminal[r][i].l[0] = 100 + r * 3 + i; }
    for (int r = 0; r < 22; r++) P.internal[r].l[0] = 1000 + r;
    Fr h[4]; memset(h, 0, sizeof h); h[0].l[0] = 5; h[1].l[0] = 7; h[2].l[0] = 9; h[3].l[0] = 11; h[3].l[5] = 77;
    Fr* io; long long* clk; cudaMalloc(&io, 64 * 32); cudaMalloc(&clk, 8);
    const char* names[] = {"fr_mul_lazy", "fr_sqr_lazy", "p2_sbox<5>", "tri_sum + add", "add + canon8", "add + reduce_once", "p2_permute_tri<5>", "fr_mul", "fr_mul (m ready early)"};
    long long c;
#define RUN(M, reps) cudaMemcpy(io, h, sizeof h, cudaMemcpyHostToDevice); k<M><<<1, 32>>>(P, io, clk, reps); k<M><<<1, 32>>>(P, io, clk, reps); \
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost); printf("%-20s %8lld cycles\n", names[M], c);
    RUN(0, 256) RUN(1, 256) RUN(2, 64) RUN(3, 256) RUN(4, 256) RUN(5, 256) RUN(6, 8) RUN(7, 256) RUN(8, 256)
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
