// Lone-warp latency probes for the multiplier's building blocks (cycles per instruction, clock64, ONE warp on the device):
// what a dependent link costs through the carry predicate, through the 64-bit accumulator, through the multiplicand, and what
// independent IMAD.WIDE issue at -- the numbers a latency-shaped Montgomery product has to be designed around.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/latprobe.cu -o tools/latprobe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 64
template <int MODE>
__global__ void __launch_bounds__(32) k(uint32_t* io, long long* clk, int reps) {
    uint32_t b = io[threadIdx.x] | 1u, x = io[32 + threadIdx.x];
    uint32_t lo[8], hi[8], cnt[8];
#pragma unroll
    for (int c = 0; c < 8; c++) { lo[c] = x * (c + 3); hi[c] = x ^ (c * 77u); cnt[c] = 0; }
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
        if (MODE == 0) {  // 64 independent IMAD.WIDE (8 accumulators, multiplicand = loop-carried b only): lone-warp issue rate
#pragma unroll
            for (int i = 0; i < N; i++) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i & 7]), "+r"(hi[i & 7]) : "r"(x + i), "r"(b));
        }
        if (MODE == 1) {  // one accumulator: every IMAD.WIDE depends on the previous through Rc (64-bit accumulate)
#pragma unroll
            for (int i = 0; i < N; i++) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(x + i), "r"(b));
        }
        if (MODE == 2) {  // linked ONLY by the carry predicate: 8 accumulators visited round-robin inside one carry chain
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
#pragma unroll
            for (int i = 0; i < N; i++) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i & 7]), "+r"(hi[i & 7]) : "r"(x + i), "r"(b));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(cnt[0]));
        }
        if (MODE == 3) {  // multiplicand dependency: each product's multiplier is the low word of the previous result
#pragma unroll
            for (int i = 0; i < N; i++) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[(i + 1) & 7]), "+r"(hi[(i + 1) & 7]) : "r"(lo[i & 7]), "r"(b));
        }
        if (MODE == 4) {  // carry-save pair: IMAD.WIDE with carry-out + IADD3.X into a counter, 8 independent accumulators
#pragma unroll
            for (int i = 0; i < N; i++)
                asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo[i & 7]), "+r"(hi[i & 7]), "+r"(cnt[i & 7]) : "r"(x + i), "r"(b));
        }
        if (MODE == 5) {  // two interleaved carry chains of 4 (the shape of one row of the current product), 8 rows
#pragma unroll
            for (int row = 0; row < 8; row++) {
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(x + row), "r"(b));
#pragma unroll
                for (int i = 1; i < 4; i++) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(x + i), "r"(b));
                asm volatile("addc.u32 %0, %0, 0;" : "+r"(cnt[0]));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[4]), "+r"(hi[4]) : "r"(x + row), "r"(b));
#pragma unroll
                for (int i = 5; i < 8; i++) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(x + i), "r"(b));
                asm volatile("addc.u32 %0, %0, 0;" : "+r"(cnt[1]));
            }
        }
        if (MODE == 6) {  // dependent IADD3 chain (ALU latency)
#pragma unroll
            for (int i = 0; i < N; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[0]) : "r"(b));
        }
        if (MODE == 7) {  // dependent IADD3.X chain through the carry
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(lo[0]) : "r"(b));
#pragma unroll
            for (int i = 1; i < N; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(lo[i & 7]) : "r"(b));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(cnt[0]));
        }
        if (MODE == 8) {  // shuffle round trip: dependent chain of SHFL
#pragma unroll
            for (int i = 0; i < N; i++) lo[0] = __shfl_xor_sync(0xffffffffu, lo[0] + 1, 1);
        }
    }
    long long t1 = clock64();
    uint32_t acc = x;
#pragma unroll
    for (int c = 0; c < 8; c++) acc ^= lo[c] ^ hi[c] ^ cnt[c];
    io[64 + threadIdx.x] = acc;
    if (threadIdx.x == 0) *clk = t1 - t0;
}

int main() {
    uint32_t h[64];
    for (int i = 0; i < 64; i++) h[i] = 0x9e3779b9u * (i + 1);
    uint32_t* io; long long* clk; cudaMalloc(&io, 128 * 4); cudaMalloc(&clk, 8);
    cudaMemcpy(io, h, sizeof h, cudaMemcpyHostToDevice);
    const char* names[] = {"independent IMAD.WIDE (8 accumulators)", "IMAD.WIDE chained through the accumulator", "IMAD.WIDE.X chained through the carry only",
                           "IMAD.WIDE chained through the multiplicand", "carry-save pair IMAD.WIDE(+P) + IADD3.X counter", "two carry chains of 4 per row (current product's shape)",
                           "IADD3 dependent chain", "IADD3.X dependent carry chain", "SHFL dependent chain"};
    long long c;
    const int reps = 64;
#define RUN(M) k<M><<<1, 32>>>(io, clk, reps); k<M><<<1, 32>>>(io, clk, reps); cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost); \
    printf("%-58s %7.2f cycles per instruction (pair)\n", names[M], double(c) / (reps * N));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
