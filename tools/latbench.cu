// Lone-warp latencies (cycles, clock64) of the pieces of the three-lanes-per-permutation shape:
// product, square, S-box, lane-triple sum, canonicalisations, and the whole permutation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I linea-stark-prover_b200/csrc tools/latbench.cu -o tools/latbench
#include <cstdio>
#include <cstring>
#include "poseidon2.cuh"
using namespace lsp;

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
    const char* names[] = {"fr_mul_lazy", "fr_sqr_lazy", "p2_sbox<5>", "tri_sum + add", "add + canon8", "add + reduce_once", "p2_permute_tri<5>", "fr_mul"};
    long long c;
#define RUN(M, reps) cudaMemcpy(io, h, sizeof h, cudaMemcpyHostToDevice); k<M><<<1, 32>>>(P, io, clk, reps); k<M><<<1, 32>>>(P, io, clk, reps); \
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost); printf("%-20s %8lld cycles\n", names[M], c);
    RUN(0, 256) RUN(1, 256) RUN(2, 64) RUN(3, 256) RUN(4, 256) RUN(5, 256) RUN(6, 8) RUN(7, 256)
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
