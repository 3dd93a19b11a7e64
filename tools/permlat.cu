// Lone-warp latency of one Poseidon2 permutation (cycles), throughput vs latency variant.
#include <cstdio>
#include "poseidon2.cuh"
using namespace lsp;
#define LSP_ERR_PARAM (-1)
template <bool LAT>
__global__ void k(const __grid_constant__ P2Params P, Fr* io, long long* clk, int reps) {
    Fr s0 = fr_load(io), s1 = fr_load(io + 1), s2 = fr_load(io + 2);
    long long t0 = clock64();
    for (int i = 0; i < reps; i++) p2_permute<5, LAT>(P, s0, s1, s2);
    long long t1 = clock64();
    fr_store(io, s0); fr_store(io + 1, s1); fr_store(io + 2, s2);
    if (threadIdx.x == 0) *clk = (t1 - t0) / reps;
}
template <bool LAT>
__global__ void kmul(Fr* io, long long* clk, int reps) {
    Fr a = fr_load(io), b = fr_load(io + 1);
    long long t0 = clock64();
    for (int i = 0; i < reps; i++) a = LAT ? fr_mul_lat_call(a, b) : fr_mul_call(a, b);
    long long t1 = clock64();
    fr_store(io, a);
    if (threadIdx.x == 0) *clk = (t1 - t0) / reps;
}
__global__ void kmul_inl(Fr* io, long long* clk, int reps) {
    Fr a = fr_load(io), b = fr_load(io + 1);
    long long t0 = clock64();
    for (int i = 0; i < reps; i++) a = fr_mul(a, b);
    long long t1 = clock64();
    fr_store(io, a);
    if (threadIdx.x == 0) *clk = (t1 - t0) / reps;
}
__global__ void kadd(Fr* io, long long* clk, int reps) {
    Fr a = fr_load(io), b = fr_load(io + 1);
    long long t0 = clock64();
    for (int i = 0; i < reps; i++) a = fr_add(a, b);
    long long t1 = clock64();
    fr_store(io, a);
    if (threadIdx.x == 0) *clk = (t1 - t0) / reps;
}
int main() {
    P2Params P; memset(&P, 0, sizeof P);
    P.half_f = 4; P.rounds_p = 22; P.sbox_d = 5; P.diag_kind = 1;
    for (int r = 0; r < 4; r++) for (int i = 0; i < 3; i++) { P.ext_initial[r][i].l[0] = r * 3 + i + 1; P.ext_terminal[r][i].l[0] = 100 + r * 3 + i; }
    for (int r = 0; r < 22; r++) P.internal[r].l[0] = 1000 + r;
    Fr h[3]; memset(h, 0, sizeof h); h[0].l[0] = 5; h[1].l[0] = 7; h[2].l[0] = 9;
    Fr* io; long long* clk; cudaMalloc(&io, 96); cudaMalloc(&clk, 8);
    long long c;
    for (int threads : {1, 32}) {
        cudaMemcpy(io, h, 96, cudaMemcpyHostToDevice);
        k<false><<<1, threads>>>(P, io, clk, 4); cudaDeviceSynchronize();
        k<false><<<1, threads>>>(P, io, clk, 8); cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        printf("threads=%2d perm chain-mul   : %lld cycles\n", threads, c);
        k<true><<<1, threads>>>(P, io, clk, 4); cudaDeviceSynchronize();
        k<true><<<1, threads>>>(P, io, clk, 8); cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        printf("threads=%2d perm f29-mul     : %lld cycles\n", threads, c);
    }
    kmul<false><<<1, 32>>>(io, clk, 64); kmul<false><<<1, 32>>>(io, clk, 256); cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("mul chain (call)  : %lld cycles\n", c);
    kmul<true><<<1, 32>>>(io, clk, 64); kmul<true><<<1, 32>>>(io, clk, 256); cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("mul f29 (call)    : %lld cycles\n", c);
    kmul_inl<<<1, 32>>>(io, clk, 64); kmul_inl<<<1, 32>>>(io, clk, 256); cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("mul chain (inline): %lld cycles\n", c);
    kadd<<<1, 32>>>(io, clk, 64); kadd<<<1, 32>>>(io, clk, 256); cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("add               : %lld cycles\n", c);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
