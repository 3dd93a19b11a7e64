// Integer-pipe microbenchmark for B200 (sm_100a): measures the issue rate of the
// instructions a 256-bit Montgomery multiply is built from.  The result
// (MAC32/s at the observed clock) is the denominator of the integer roofline
// reported by bench.py; written to gpurun_out/int_peak.json.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 17u + seed;
    uint32_t lo[CHAINS], hi[CHAINS];
    double d[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { lo[c] = a + c; hi[c] = b ^ c; d[c] = (double)(a + c); }
    double da = (double)a * 1e-9, db = (double)b * 1e-9;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (MODE == 0) {  // mad.lo.u32 (IMAD)
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(a), "r"(b));
            } else if (MODE == 1) {  // mad.hi.u32
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(a), "r"(b));
            } else if (MODE == 2) {  // mad.wide.u32 (IMAD.WIDE.U32), independent 64-bit accumulators
                asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %2, %3, t; mov.b64 {%0,%1}, t;}"
                             : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b));
            } else if (MODE == 5) {  // IMAD.WIDE + IADD3 pair (does the ALU pipe co-issue?)
                asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %2, %3, t; mov.b64 {%0,%1}, t;}"
                             : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b));
            } else if (MODE == 6) {  // DFMA
                asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[c]) : "d"(da), "d"(db));
            }
        }
        if (MODE == 3) {  // one carry chain across all CHAINS: lo/hi pairs with cc
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(a), "r"(b));
#pragma unroll
            for (int c = 1; c < CHAINS; c++)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b));
        }
        if (MODE == 4) {  // two independent carry chains of CHAINS/2
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int o = h * (CHAINS / 2);
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[o]), "+r"(hi[o]) : "r"(a), "r"(b));
#pragma unroll
                for (int c = 1; c < CHAINS / 2; c++)
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[o + c]), "+r"(hi[o + c]) : "r"(a), "r"(b));
            }
        }
        if (MODE == 5) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(lo[c]));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) r ^= lo[c] ^ hi[c] ^ (uint32_t)d[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r ^ a;
}

template <int MODE>
double run(const char* name, double ops_per_iter_per_thread, uint32_t* out, FILE* js, bool last) {
    int blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) k<MODE><<<blocks, threads>>>(out, w);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(out, r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)blocks * threads * ITERS * ops_per_iter_per_thread;
    double rate = ops / (best * 1e-3);
    printf("%-28s %8.3f ms  %8.3f Tops/s  (%.1f per clk per SM @1.965GHz)\n", name, best, rate / 1e12, rate / 148 / 1.965e9);
    fprintf(js, "  \"%s\": %.6e%s\n", name, rate, last ? "" : ",");
    return rate;
}

int main() {
    uint32_t* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    FILE* js = fopen("gpurun_out/int_peak.json", "w");
    if (!js) js = stdout;
    fprintf(js, "{\n");
    run<0>("imad_lo", CHAINS, out, js, false);
    run<1>("imad_hi", CHAINS, out, js, false);
    run<2>("imad_wide", CHAINS, out, js, false);
    run<3>("wide_cc_chain8", CHAINS, out, js, false);
    run<4>("wide_cc_chain4x2", CHAINS, out, js, false);
    run<5>("imad_wide_plus_iadd", CHAINS, out, js, false);
    run<6>("dfma", CHAINS, out, js, true);
    fprintf(js, "}\n");
    if (js != stdout) fclose(js);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
