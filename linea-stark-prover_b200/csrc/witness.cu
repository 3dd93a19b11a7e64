// Witness generation for the permutation argument on the device.
//
// Replaces `RawPermutationTrace::get_trace` (reference trace/src/permutation.rs:24-93)
// followed by `RawTrace::get_trace` (trace/src/lib.rs:94-106):
//   columns  a.., b.., 1/(b_comb + delta), running product of (a_comb + delta)/(b_comb + delta)
// where x_comb is the Horner combination of the row in alpha (:56-68).
// The reference inverts once per row (:70) and multiplies sequentially (:72);
// here the inversions are batched with Montgomery's trick (one Fermat inversion
// per 8 rows per thread) and the running product is a three-phase parallel scan.
// The last-row assertion (:76-79) is kept: a non-permutation is an error.
#include "stark.cuh"

using namespace lsp;

namespace {

constexpr int WIT_BATCH = 8;

// in: host-layout row-major N x 2c (a columns then b columns).  out: column-major N x (2c+2).
// Writes a/b columns, den = b_comb + delta into column 2c, num = a_comb + delta into column 2c+1.
__global__ void __launch_bounds__(128) k_wit_combine(const Fr* __restrict__ ab_rm, size_t n, int c, const Fr* __restrict__ publics,
                                                     Fr* __restrict__ out) {
    const Fr alpha = fr_load(publics), delta = fr_load(publics + 1);
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const Fr* row = ab_rm + i * size_t(2 * c);
        Fr ac = fr_zero(), bc = fr_zero();
        for (int j = 0; j < c; j++) {
            Fr a = fr_load_nc(row + j), b = fr_load_nc(row + c + j);
            fr_store(out + size_t(j) * n + i, a);
            fr_store(out + size_t(c + j) * n + i, b);
            ac = j ? fr_add(fr_mul(ac, alpha), a) : a;
            bc = j ? fr_add(fr_mul(bc, alpha), b) : b;
        }
        fr_store(out + size_t(2 * c) * n + i, fr_add(bc, delta));
        fr_store(out + size_t(2 * c + 1) * n + i, fr_add(ac, delta));
    }
}

// den -> 1/den in place (column 2c) and num -> num/den (column 2c+1).  Each thread owns
// WIT_BATCH elements spaced by the thread count (coalesced) and does one inversion.
__global__ void __launch_bounds__(128) k_wit_invert(Fr* __restrict__ den, Fr* __restrict__ num, size_t n) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t base = blockIdx.x * size_t(blockDim.x) + threadIdx.x; base < n; base += stride * WIT_BATCH) {
        Fr x[WIT_BATCH], pre[WIT_BATCH];
        Fr acc = fr_one();
        int m = 0;
#pragma unroll
        for (int k = 0; k < WIT_BATCH; k++) {
            size_t i = base + k * stride;
            if (i < n) {
                x[k] = fr_load(den + i);
                pre[k] = acc;
                acc = fr_mul(acc, x[k]);
                m = k + 1;
            }
        }
        acc = fr_inv(acc);
#pragma unroll
        for (int k = WIT_BATCH - 1; k >= 0; k--) {
            if (k < m) {
                size_t i = base + k * stride;
                Fr inv = fr_mul(acc, pre[k]);
                acc = fr_mul(acc, x[k]);
                fr_store(den + i, inv);
                fr_store(num + i, fr_mul(fr_load(num + i), inv));
            }
        }
    }
}

// ---- inclusive prefix product over n elements: tile = 1024 elements per block -----------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ void sm_put(uint4* s, int i, const Fr& v) {
    s[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    s[SCAN_THREADS + i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr sm_get(const uint4* s, int i) {
    uint4 a = s[i], b = s[SCAN_THREADS + i];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// Phase 1: in-tile inclusive scan (in place) + tile totals.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(Fr* __restrict__ v, size_t n, Fr* __restrict__ totals) {
    __shared__ uint4 sm[2 * SCAN_THREADS];
    const size_t t0 = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_PER_THREAD;
    Fr x[SCAN_PER_THREAD];
    Fr acc = fr_one();
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        x[k] = (t0 + k < n) ? fr_load(v + t0 + k) : fr_one();
        acc = fr_mul(acc, x[k]);
        x[k] = acc;
    }
    sm_put(sm, threadIdx.x, acc);
    __syncthreads();
    for (int off = 1; off < SCAN_THREADS; off <<= 1) {  // Hillis-Steele over the per-thread totals
        Fr mine = sm_get(sm, threadIdx.x), other = fr_one();
        bool has = int(threadIdx.x) >= off;
        if (has) other = sm_get(sm, threadIdx.x - off);
        __syncthreads();
        if (has) sm_put(sm, threadIdx.x, fr_mul(other, mine));
        __syncthreads();
    }
    Fr before = threadIdx.x ? sm_get(sm, threadIdx.x - 1) : fr_one();
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++)
        if (t0 + k < n) fr_store(v + t0 + k, threadIdx.x ? fr_mul(before, x[k]) : x[k]);
    if (threadIdx.x == SCAN_THREADS - 1) fr_store(totals + blockIdx.x, sm_get(sm, SCAN_THREADS - 1));
}
// Phase 2: exclusive scan of the tile totals by one block (sequential over chunks of 256).
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_totals(Fr* __restrict__ totals, size_t n_tiles) {
    __shared__ uint4 sm[2 * SCAN_THREADS];
    Fr carry = fr_one();
    for (size_t base = 0; base < n_tiles; base += SCAN_THREADS) {
        size_t i = base + threadIdx.x;
        Fr mine = i < n_tiles ? fr_load(totals + i) : fr_one();
        sm_put(sm, threadIdx.x, mine);
        __syncthreads();
        for (int off = 1; off < SCAN_THREADS; off <<= 1) {
            Fr cur = sm_get(sm, threadIdx.x), other = fr_one();
            bool has = int(threadIdx.x) >= off;
            if (has) other = sm_get(sm, threadIdx.x - off);
            __syncthreads();
            if (has) sm_put(sm, threadIdx.x, fr_mul(other, cur));
            __syncthreads();
        }
        Fr excl = threadIdx.x ? sm_get(sm, threadIdx.x - 1) : fr_one();
        Fr last = sm_get(sm, SCAN_THREADS - 1);
        if (i < n_tiles) fr_store(totals + i, fr_mul(carry, excl));
        carry = fr_mul(carry, last);
        __syncthreads();
    }
}
// Phase 3: multiply every tile by the product of the tiles before it.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(Fr* __restrict__ v, size_t n, const Fr* __restrict__ totals) {
    if (blockIdx.x == 0) return;
    Fr pre = fr_load(totals + blockIdx.x);
    const size_t t0 = size_t(blockIdx.x) * SCAN_TILE;
    for (int k = threadIdx.x; k < SCAN_TILE; k += SCAN_THREADS)
        if (t0 + k < n) fr_store(v + t0 + k, fr_mul(pre, fr_load(v + t0 + k)));
}

__global__ void k_check_last_is_one(const Fr* __restrict__ v, size_t n, int* __restrict__ flag) {
    *flag = fr_eq(fr_load(v + n - 1), fr_one()) ? 0 : 1;
}

}  // namespace

// `from_be_bytes_mod_order` + Montgomery conversion, in place: 32 big-endian bytes -> x*R mod r.
// x < 2^256 goes in as the unrestricted operand of the lazy product with R^2 (< r).
__global__ void __launch_bounds__(128) k_be_to_mont(Fr* __restrict__ v, size_t count) {
    const Fr r2 = fr_const(FR_R2);
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x) {
        Fr be = fr_load(v + i), x;
#pragma unroll
        for (int k = 0; k < 8; k++) x.l[k] = __byte_perm(be.l[7 - k], 0, 0x0123);
        Fr y = fr_mul_lazy(r2, x);
        fr_reduce_once(y);
        fr_store(v + i, y);
    }
}

static int permutation_trace_impl(lsp_ctx* ctx, const void* ab_rowmajor, bool big_endian_bytes, size_t rows, uint32_t n_cols,
                                  const uint64_t publics[2][4], lsp_mat** trace_out);

extern "C" int lsp_permutation_trace(lsp_ctx* ctx, const uint64_t* ab_rowmajor, size_t rows, uint32_t n_cols,
                                     const uint64_t publics[2][4], lsp_mat** trace_out) {
    return permutation_trace_impl(ctx, ab_rowmajor, false, rows, n_cols, publics, trace_out);
}

extern "C" int lsp_permutation_trace_be(lsp_ctx* ctx, const uint8_t* be_rowmajor, size_t rows, uint32_t n_cols,
                                        const uint64_t publics[2][4], lsp_mat** trace_out) {
    return permutation_trace_impl(ctx, be_rowmajor, true, rows, n_cols, publics, trace_out);
}

static int permutation_trace_impl(lsp_ctx* ctx, const void* ab_rowmajor, bool big_endian_bytes, size_t rows, uint32_t n_cols,
                                  const uint64_t publics[2][4], lsp_mat** trace_out) {
    if (!ctx || !ab_rowmajor || !publics || !trace_out || rows == 0 || n_cols == 0) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = rows, w_in = 2 * size_t(n_cols), w_out = w_in + 2;
    Fr *stage = nullptr, *pub = nullptr, *totals = nullptr;
    int* flag = nullptr;
    const size_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    LSP_TRY(dev_alloc(ctx, (void**)&stage, n * w_in * 32));
    LSP_TRY(dev_alloc(ctx, (void**)&pub, 64));
    LSP_TRY(dev_alloc(ctx, (void**)&totals, n_tiles * 32));
    LSP_TRY(dev_alloc(ctx, (void**)&flag, 4));
    lsp_mat* m = nullptr;
    LSP_TRY(mat_alloc(ctx, n, w_out, &m));
    LSP_CUDA(ctx, cudaMemcpyAsync(stage, ab_rowmajor, n * w_in * 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemcpyAsync(pub, publics, 64, cudaMemcpyHostToDevice, ctx->stream));
    if (big_endian_bytes) LSP_LAUNCH(ctx, k_be_to_mont, grid_for(ctx, n * w_in, 128), 128, 0, stage, n * w_in);
    Fr* den = m->d + w_in * n;
    Fr* num = den + n;
    LSP_LAUNCH(ctx, k_wit_combine, grid_for(ctx, n, 128), 128, 0, (const Fr*)stage, n, int(n_cols), (const Fr*)pub, m->d);
    LSP_LAUNCH(ctx, k_wit_invert, grid_for(ctx, (n + WIT_BATCH - 1) / WIT_BATCH, 128), 128, 0, den, num, n);
    LSP_LAUNCH(ctx, k_scan_tiles, unsigned(n_tiles), SCAN_THREADS, 0, num, n, totals);
    LSP_LAUNCH(ctx, k_scan_totals, 1, SCAN_THREADS, 0, totals, n_tiles);
    LSP_LAUNCH(ctx, k_scan_apply, unsigned(n_tiles), SCAN_THREADS, 0, num, n, (const Fr*)totals);
    LSP_LAUNCH(ctx, k_check_last_is_one, 1, 1, 0, (const Fr*)num, n, flag);
    LSP_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, stage);
    dev_free(ctx, pub);
    dev_free(ctx, totals);
    dev_free(ctx, flag);
    if (*(volatile int*)ctx->pinned) {
        lsp_mat_free(ctx, m);
        return set_err(ctx, LSP_ERR_PARAM, "failed to check constrain: check column should be 1 on the last row");
    }
    *trace_out = m;
    return LSP_OK;
}
