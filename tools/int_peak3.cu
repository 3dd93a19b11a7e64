// Integer-pipe microbenchmark v3 (sm_100a): data-dependent multiplicands so ptxas cannot
// strength-reduce; reports issue cycles per warp-instruction per SMSP at the observed clock.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
typedef unsigned long long u64;

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, long long* clk) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 17u + seed;
    u64 acc[8];
    uint32_t x[8], y[8];
    double d[8];
#pragma unroll
    for (int c = 0; c < 8; c++) { acc[c] = ((u64)(b ^ c) << 32) | (a + c); x[c] = a * (c + 3); y[c] = a * (c + 5) + b; d[c] = a + c; }
    double da = a * 1e-9, db = b * 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 5 || MODE == 6 || MODE == 7 || MODE == 10) {  // plain IMAD.WIDE
#pragma unroll
            for (int c = 0; c < 8; c++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"((uint32_t)acc[(c + 1) & 7]), "r"(b));
        }
        if (MODE == 1) {  // IMAD.WIDE with carry chain (.X)
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(a));
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t lo = (uint32_t)acc[c], hi = (uint32_t)(acc[c] >> 32);
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"((uint32_t)acc[(c + 1) & 7]), "r"(b));
                acc[c] = ((u64)hi << 32) | lo;
            }
        }
        if (MODE == 2 || MODE == 8) {  // IMAD.LO 32-bit
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[c]) : "r"(x[(c + 1) & 7]), "r"(b));
        }
        if (MODE == 9) {  // IMAD.HI 32-bit
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x[c]) : "r"(x[(c + 1) & 7]), "r"(b));
        }
        if (MODE == 3 || MODE == 8) {  // + 8 IADD3 (3-input adds)
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(y[c]) : "r"(y[(c + 1) & 7]), "r"(y[(c + 3) & 7]));
        }
        if (MODE == 4) {  // + 8 LOP3
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[c]) : "r"(y[(c + 1) & 7]), "r"(y[(c + 3) & 7]));
        }
        if (MODE == 5) {  // + 8 SHF
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("shf.r.clamp.b32 %0, %0, %1, 29;" : "+r"(y[c]) : "r"(y[(c + 1) & 7]));
        }
        if (MODE == 6) {  // + 8 DFMA
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(da), "d"(db));
        }
        if (MODE == 7) {  // + 8 add.cc chain (IADD3.X)
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(y[0]) : "r"(y[1]));
#pragma unroll
            for (int c = 1; c < 8; c++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(y[c]) : "r"(y[(c + 1) & 7]));
        }
        if (MODE == 10) {  // + 16 IADD3
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(y[c]) : "r"(y[(c + 1) & 7]), "r"(y[(c + 3) & 7]));
        }
        if (MODE == 11) {  // 8 IADD3 only
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(y[c]) : "r"(y[(c + 1) & 7]), "r"(y[(c + 3) & 7]));
        }
        if (MODE == 12) {  // 8 LOP3 only
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[c]) : "r"(y[(c + 1) & 7]), "r"(y[(c + 3) & 7]));
        }
        if (MODE == 13) {  // 8 DFMA only
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(da), "d"(db));
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) r ^= (uint32_t)acc[c] ^ (uint32_t)(acc[c] >> 32) ^ x[c] ^ y[c] ^ (uint32_t)d[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int MODE>
void run(const char* name, double n_instr, uint32_t* out, long long* clk) {
    int blocks = 148 * 8, threads = 256;   // 8 blocks x 8 warps = 64 warps per SM, 16 per SMSP
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) k<MODE><<<blocks, threads>>>(out, w, clk);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(out, r, clk);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    // SM cycles per warp-instruction per SMSP from the in-kernel clock (clock independent):
    // block 0 ran with 16 warps per SMSP resident
    double cyc = (double)h / (ITERS * n_instr * 16.0);
    printf("%-36s %8.3f ms  %6.2f cyc/warp-instr/SMSP (clock64)  [%g instr/iter]\n", name, best, cyc, n_instr);
}

int main() {
    uint32_t* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    long long* clk; cudaMalloc(&clk, 8);
    run<0>("8 IMAD.WIDE", 8, out, clk);
    run<1>("8 IMAD.WIDE.X (carry chain)", 8, out, clk);
    run<2>("8 IMAD.LO", 8, out, clk);
    run<9>("8 IMAD.HI", 8, out, clk);
    run<11>("8 IADD3", 8, out, clk);
    run<12>("8 LOP3", 8, out, clk);
    run<13>("8 DFMA", 8, out, clk);
    run<3>("8 IMAD.WIDE + 8 IADD3", 16, out, clk);
    run<10>("8 IMAD.WIDE + 16 IADD3", 24, out, clk);
    run<4>("8 IMAD.WIDE + 8 LOP3", 16, out, clk);
    run<5>("8 IMAD.WIDE + 8 SHF", 16, out, clk);
    run<6>("8 IMAD.WIDE + 8 DFMA", 16, out, clk);
    run<7>("8 IMAD.WIDE + 8 IADD3.X", 16, out, clk);
    run<8>("8 IMAD.LO + 8 IADD3", 16, out, clk);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
