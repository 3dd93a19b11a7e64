// Second integer-pipe microbenchmark: which instruction mixes issue at full rate on sm_100a?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096

// MODE 0: IMAD.WIDE plain x8
// MODE 1: IMAD.WIDE carry-out only (mad.lo.cc/madc.hi without carry-in) x8
// MODE 2: 4 plain + 4 .X (chain of 4 after a cc starter)  -- mixed heavy/lite?
// MODE 3: 8 plain IMAD.WIDE + 8 independent IADD3 (co-issue FMA+ALU)
// MODE 4: 8 IADD3 only
// MODE 5: 8 IADD3.X chains (add.cc/addc.cc)
// MODE 6: 8 SHF (funnel shift) only
// MODE 7: 8 LOP3 only
// MODE 8: 8 plain IMAD.WIDE + 16 independent IADD3
// MODE 9: 8 IMAD.WIDE.X + 8 IADD3
// MODE 10: 8 plain IMAD.WIDE + 8 DFMA
// MODE 11: 8 IMAD.LO (32-bit) + 8 plain IMAD.WIDE
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 17u + seed;
    uint32_t lo[8], hi[8], x[16];
    double d[8];
#pragma unroll
    for (int c = 0; c < 8; c++) { lo[c] = a + c; hi[c] = b ^ c; d[c] = a + c; }
#pragma unroll
    for (int c = 0; c < 16; c++) x[c] = a * (c + 3);
    double da = a * 1e-9, db = b * 1e-9;
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0 || MODE == 3 || MODE == 8 || MODE == 10 || MODE == 11) {
#pragma unroll
            for (int c = 0; c < 8; c++)
                asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %0, %3, t; mov.b64 {%0,%1}, t;}" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b));
        }
        if (MODE == 1) {
#pragma unroll
            for (int c = 0; c < 8; c++)
                asm volatile("mad.lo.cc.u32 %0, %1, %3, %0; madc.hi.u32 %1, %1, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b));
        }
        if (MODE == 2) {
#pragma unroll
            for (int c = 0; c < 4; c++)
                asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %0, %3, t; mov.b64 {%0,%1}, t;}" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b));
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(a));
#pragma unroll
            for (int c = 4; c < 8; c++)
                asm volatile("madc.lo.cc.u32 %0, %1, %3, %0; madc.hi.cc.u32 %1, %1, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b));
        }
        if (MODE == 9) {
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[15]) : "r"(a));
#pragma unroll
            for (int c = 0; c < 8; c++)
                asm volatile("madc.lo.cc.u32 %0, %1, %3, %0; madc.hi.cc.u32 %1, %1, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b));
        }
        if (MODE == 3 || MODE == 4 || MODE == 9) {
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(x[(c+1)&15]));
        }
        if (MODE == 8) {
#pragma unroll
            for (int c = 0; c < 16; c++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(x[(c+1)&15]));
        }
        if (MODE == 5) {
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(b));
#pragma unroll
            for (int c = 1; c < 8; c++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(b));
        }
        if (MODE == 6) {
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("shf.r.clamp.b32 %0, %0, %1, 29;" : "+r"(x[c]) : "r"(b));
        }
        if (MODE == 7) {
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(b), "r"(a));
        }
        if (MODE == 10) {
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[c]) : "d"(da), "d"(db));
        }
        if (MODE == 11) {
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("mad.lo.u32 %0, %0, %2, %0;" : "+r"(x[c]) : "r"(a), "r"(b));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) r ^= lo[c] ^ hi[c] ^ (uint32_t)d[c];
#pragma unroll
    for (int c = 0; c < 16; c++) r ^= x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double n_instr, uint32_t* out) {
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
    double warp_instr = (double)blocks * threads / 32 * ITERS * n_instr;
    // cycles per warp-instruction per SMSP assuming 1.965 GHz
    double cyc = best * 1e-3 * 1.965e9 * 148 * 4 / warp_instr;
    printf("%-34s %8.3f ms   %.2f issue-cycles per warp-instr per SMSP (n_instr=%g)\n", name, best, cyc, n_instr);
}

int main() {
    uint32_t* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("8 imad.wide", 8, out);
    run<1>("8 imad.wide carry-out only", 8, out);
    run<2>("4 imad.wide + 4 imad.wide.X", 9, out);
    run<3>("8 imad.wide + 8 iadd3", 16, out);
    run<4>("8 iadd3", 8, out);
    run<5>("8 iadd3.X chain", 8, out);
    run<6>("8 shf", 8, out);
    run<7>("8 lop3", 8, out);
    run<8>("8 imad.wide + 16 iadd3", 24, out);
    run<9>("8 imad.wide.X + 8 iadd3", 17, out);
    run<10>("8 imad.wide + 8 dfma", 16, out);
    run<11>("8 imad.wide + 8 imad.lo", 16, out);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
