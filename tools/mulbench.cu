// Field-multiplication shoot-out on sm_100a: modmul/s of candidate Montgomery products,
// each checked against the 32-bit carry-chain product that the parity tests pin to the oracle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I linea-stark-prover_b200/csrc tools/mulbench.cu -o tools/mulbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fr.cuh"
#include "fr29.cuh"
#include "fr_mul_v1.cuh"
#include "fr_mul_split.cuh"

using namespace lsp;
#define CHAINS 2
#define ITERS 512

// MODE 0: fr_mul_v1 (round-1a product)  1: fr_mul  2: f29 packed in/out per mul  3: f29 persistent
// MODE 4: x = x*x with fr_mul_v1  5: x = fr_sqr(x)  6: x = y*x then x = x^2 (S-box-like mix)
template <int MODE>
__global__ void __launch_bounds__(128) k(const Fr* __restrict__ in, Fr* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    Fr x[CHAINS], y[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { x[c] = fr_load(in + (size_t(i) * CHAINS + c) % n); y[c] = fr_load(in + (size_t(i) * CHAINS + c + 7) % n); }
    if (MODE == 3) {
        F29 a[CHAINS], b[CHAINS];
#pragma unroll
        for (int c = 0; c < CHAINS; c++) { a[c] = f29_unpack(x[c]); b[c] = f29_unpack(y[c]); }
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) a[c] = f29_mul(a[c], b[c]);
        }
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = f29_pack_canonical(a[c]);
    } else {
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                if (MODE == 0) x[c] = lsp_v1::fr_mul_v1(x[c], y[c]);
                if (MODE == 1) x[c] = fr_mul(x[c], y[c]);
                if (MODE == 4) x[c] = lsp_v1::fr_mul_v1(x[c], x[c]);
                if (MODE == 5) x[c] = fr_sqr(x[c]);
                if (MODE == 7) x[c] = lsp_split::fr_mul_split(x[c], y[c]);
                if (MODE == 2) x[c] = f29_pack_lazy(f29_mul(f29_unpack(x[c]), f29_unpack(y[c])));
            }
        }
        if (MODE == 2) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) fr_reduce_once(x[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < CHAINS; c++) fr_store(out + size_t(i) * CHAINS + c, x[c]);
}

template <int MODE>
float run(const char* name, const Fr* in, Fr* out, int n, int blocks) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) k<MODE><<<blocks, 128>>>(in, out, n);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 128>>>(in, out, n);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double muls = double(blocks) * 128 * CHAINS * ITERS;
    printf("%-28s %8.3f ms  %7.2f G modmul/s\n", name, best, muls / (best * 1e-3) / 1e9);
    return best;
}

int main() {
    const int n = 1 << 16, blocks = 148 * 16;
    Fr* h = (Fr*)malloc(n * sizeof(Fr));
    uint64_t s = 88172645463325252ull;
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < 8; j++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i].l[j] = (uint32_t)s; }
        h[i].l[7] &= 0x0fffffffu;  // < 2^252 < r
    }
    Fr *in, *o0, *o1;
    size_t out_n = size_t(blocks) * 128 * CHAINS;
    cudaMalloc(&in, n * sizeof(Fr)); cudaMalloc(&o0, out_n * sizeof(Fr)); cudaMalloc(&o1, out_n * sizeof(Fr));
    cudaMemcpy(in, h, n * sizeof(Fr), cudaMemcpyHostToDevice);
    run<0>("fr_mul_v1 (round 1a)", in, o0, n, blocks);
    run<1>("fr_mul (120 IMAD.WIDE)", in, o1, n, blocks);
    Fr* a = (Fr*)malloc(out_n * sizeof(Fr)); Fr* b = (Fr*)malloc(out_n * sizeof(Fr));
    cudaMemcpy(a, o0, out_n * sizeof(Fr), cudaMemcpyDeviceToHost);
    auto check = [&](const char* nm, Fr* dev) {
        cudaMemcpy(b, dev, out_n * sizeof(Fr), cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t i = 0; i < out_n; i++) if (memcmp(&a[i], &b[i], 32)) bad++;
        printf("   %s vs v1: %zu mismatches of %zu\n", nm, bad, out_n);
    };
    check("fr_mul", o1);
    run<2>("f29 unpack/mul/pack", in, o1, n, blocks);
    check("f29 packed", o1);
    run<3>("f29 persistent", in, o1, n, blocks);
    check("f29 persistent", o1);
    run<0>("fr_mul_v1 (again, reference)", in, o0, n, blocks);
    cudaMemcpy(a, o0, out_n * sizeof(Fr), cudaMemcpyDeviceToHost);
    run<7>("fr_mul_split (mul-only + adds)", in, o1, n, blocks);
    check("fr_mul_split", o1);
    run<4>("v1 x*x chain", in, o0, n, blocks);
    cudaMemcpy(a, o0, out_n * sizeof(Fr), cudaMemcpyDeviceToHost);
    run<5>("fr_sqr (92 IMAD.WIDE)", in, o1, n, blocks);
    check("fr_sqr", o1);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
