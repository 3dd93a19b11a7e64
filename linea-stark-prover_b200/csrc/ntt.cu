// Radix-2 NTT passes staged in shared memory and the coset LDE built from them.
//
// Replaces `Radix2DitParallel<Val>::coset_lde_batch` (reference Dft alias,
// bin/src/config.rs:22; consumed by TwoAdicFriPcs::commit, bin/src/main.rs:66).
//
// Layout: matrices are column-major on the device, so every column is one
// contiguous polynomial and all columns of a call are batched over gridDim.y.
//   inverse  : DIT, bit-reversed gather of the evaluations fused into the first
//              pass' loads, natural-order coefficients out, 1/N fused into the
//              last pass' stores;
//   forward  : DIF on each of the 2^added_bits cosets (gridDim.z), the coset
//              powers (shift*w_L^c)^k fused into the first pass' loads, and the
//              bit-reversed row order of the LDE falls out of DIF for free:
//              coset c lands in row block bitrev(c), position bitrev(k).
// A pass runs up to LSP_NTT_MAX_T butterfly stages on a tile of 2^t elements held
// in shared memory as two 16-byte planes (conflict-free 128-bit accesses), two stages per barrier in registers, with
// the tile's twiddles and coset factors staged in shared memory once and reused by every column (k_ntt_tile).
#include "stark.cuh"

using namespace lsp;

namespace lsp {

constexpr int NTT_MAX_T = 10;  // 2^10 elements = 32 KiB of shared memory per tile

// tw[j] = w^j (or w^-j) for j < n/2, w = omega_{2^log_n}
__global__ void __launch_bounds__(128) k_gen_twiddles(const FieldConsts* __restrict__ fc, Fr* __restrict__ tw, int log_n, int inverse) {
    size_t half = (size_t(1) << log_n) >> 1;
    Fr w = fr_two_adic_generator(fc, log_n);
    for (size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x; j < half; j += size_t(gridDim.x) * blockDim.x) {
        uint32_t e = inverse ? uint32_t(((size_t(1) << log_n) - j) & ((size_t(1) << log_n) - 1)) : uint32_t(j);
        fr_store(tw + j, fr_pow_u32(w, e));
    }
}

// Two-level power tables of the coset bases S_c = shift * w_L^c, c < 2^added_bits:
//   lo[c][j] = S_c^j            j < 2^lo_bits
//   hi[c][j] = S_c^(j<<lo_bits) j < 2^(log_n - lo_bits)
__global__ void __launch_bounds__(128) k_coset_pow_tables(const FieldConsts* __restrict__ fc, Fr* __restrict__ lo, Fr* __restrict__ hi, const Fr* __restrict__ shift, int log_n,
                                                          int added_bits, int lo_bits, int block0) {
    // destination row block (block0 + blockIdx.y) of the bit-reversed LDE holds coset bitrev(block)
    int c = blockIdx.y;
    int log_l = log_n + added_bits;
    uint32_t coset = bitrev32(uint32_t(block0 + c), added_bits);
    Fr base = fr_mul(fr_load(shift), fr_pow_u32(fr_two_adic_generator(fc, log_l), coset));
    size_t n_lo = size_t(1) << lo_bits, n_hi = size_t(1) << (log_n - lo_bits);
    for (size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x; j < n_lo + n_hi; j += size_t(gridDim.x) * blockDim.x) {
        if (j < n_lo)
            fr_store(lo + c * n_lo + j, fr_pow_u32(base, uint32_t(j)));
        else
            fr_store(hi + c * n_hi + (j - n_lo), fr_pow_u32(base, uint32_t((j - n_lo) << lo_bits)));
    }
}

struct NttPass {
    const Fr* src;      // column c of the source at src + c*src_col_stride (+ z*src_z_stride)
    Fr* dst;
    size_t src_col_stride, dst_col_stride;
    size_t src_z_stride, dst_z_stride;  // per coset (gridDim.z)
    const Fr* tw;       // n/2 twiddles of the size-n transform
    int log_n;          // transform size
    int bit_lo, t;      // this pass handles index bits [bit_lo, bit_lo+t)
    int bitrev_load;    // gather src at bitrev_{log_n}(index)
    const Fr* pow_lo;   // optional coset power tables (per z)
    const Fr* pow_hi;
    int pow_lo_bits;
    int scale_store;    // multiply by `scale` on store (1/N)
    Fr scale;
    int width;          // columns of this launch (k_ntt_tile loops over them)
    int cols_per_block;
    int n_z;            // cosets of this launch
};

__device__ __forceinline__ void smem_put(uint4* plo, uint4* phi, int i, const Fr& v) {
    plo[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    phi[i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr smem_get(const uint4* plo, const uint4* phi, int i) {
    uint4 a = plo[i], b = phi[i];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// ---- the pass kernel the LDE runs on: register-blocked, column-looped --------------------------------------------
// One block = one tile of 2^t elements (one coset), LOOPED over up to `cols_per_block` columns:
//   * the twiddles the tile's t stages need (2^t - 1 of them: stage with local bit lb uses 2^lb) are fetched ONCE into
//     shared memory and reused by every column, instead of one global load per butterfly;
//   * the coset power S_c^g of every element (two table entries multiplied) is computed ONCE per tile and kept in shared
//     memory: one product per element per column instead of two;
//   * a thread owns FOUR elements and runs TWO butterfly stages on them in registers between barriers (a radix-4 step
//     costs 4 products and 3 twiddle reads; in a prime field there is no cheaper radix-4 butterfly): half the barriers
//     and half the shared-memory traffic of one stage per barrier, and two independent products in flight per thread.
// Shared memory: tile + twiddles (+ factors) = 2 or 3 planes of 2^t x 32 bytes (96 KiB at t = 10 with factors).
#ifndef LSP_NTT_MINB
#define LSP_NTT_MINB 2
#endif
constexpr int NTT_TILE_MINB = LSP_NTT_MINB;   // resident blocks per SM the kernel is compiled for (<= 128 registers at 256 threads)

template <bool DIF>
__device__ __forceinline__ void ntt_bfly(Fr& u, Fr& v, const Fr& w, bool trivial) {
    if (DIF) {
        const Fr d = fr_sub(u, v);
        u = fr_add(u, v);
        v = trivial ? d : fr_mul(d, w);
    } else {
        const Fr vw = trivial ? v : fr_mul(v, w);
        v = fr_sub(u, vw);
        u = fr_add(u, vw);
    }
}

template <bool DIF>
__global__ void __launch_bounds__(256, NTT_TILE_MINB) k_ntt_tile(const __grid_constant__ NttPass P) {
    extern __shared__ uint4 smem[];
    const int t = P.t, tile = 1 << t;
    uint4 *plo = smem, *phi = smem + tile;               // the tile
    uint4 *wlo = smem + 2 * tile, *whi = smem + 3 * tile;  // twiddles: stage with local bit lb at [2^lb - 1, 2^(lb+1) - 1)
    uint4 *flo = smem + 4 * tile, *fhi = smem + 5 * tile;  // coset factors (first forward pass only)
    const size_t lo_mask = (size_t(1) << P.bit_lo) - 1;
    // blockIdx.x = tile * n_z + coset: the cosets of one tile run side by side, so the coefficient tile they all read
    // comes out of L2 for all but the first of them
    const int z = int(blockIdx.x % unsigned(P.n_z));
    const size_t tile_id = blockIdx.x / unsigned(P.n_z);
    const size_t idx_lo = tile_id & lo_mask, idx_hi = tile_id >> P.bit_lo;
    const size_t base = (idx_hi << (P.bit_lo + t)) | idx_lo;  // + j << bit_lo

    for (int k = threadIdx.x; k < tile - 1; k += blockDim.x) {
        const int lb = 31 - __clz(k + 1), u = k + 1 - (1 << lb), b = P.bit_lo + lb;
        const size_t e = (idx_lo + (size_t(u) << P.bit_lo)) << (P.log_n - 1 - b);
        smem_put(wlo, whi, k, fr_load_nc(P.tw + e));
    }
    if (P.pow_lo) {
        const size_t n_lo = size_t(1) << P.pow_lo_bits, n_hi = size_t(1) << (P.log_n - P.pow_lo_bits);
        for (int j = threadIdx.x; j < tile; j += blockDim.x) {
            const size_t g = base + (size_t(j) << P.bit_lo);
            smem_put(flo, fhi, j, fr_mul(fr_load_nc(P.pow_lo + z * n_lo + (g & (n_lo - 1))), fr_load_nc(P.pow_hi + z * n_hi + (g >> P.pow_lo_bits))));
        }
    }
    const int c0 = blockIdx.y * P.cols_per_block;
    const int c1 = c0 + P.cols_per_block < P.width ? c0 + P.cols_per_block : P.width;
    for (int c = c0; c < c1; c++) {
        const Fr* src = P.src + size_t(c) * P.src_col_stride + z * P.src_z_stride;
        Fr* dst = P.dst + size_t(c) * P.dst_col_stride + z * P.dst_z_stride;
        __syncthreads();   // the previous column's stores have read the tile; twiddles / factors are in place
        for (int j = threadIdx.x; j < tile; j += blockDim.x) {
            const size_t g = base + (size_t(j) << P.bit_lo);
            const size_t sidx = P.bitrev_load ? size_t(bitrev32(uint32_t(g), P.log_n)) : g;
            Fr v = fr_load_nc(src + sidx);
            if (P.pow_lo) v = fr_mul(v, smem_get(flo, fhi, j));
            smem_put(plo, phi, j, v);
        }
        __syncthreads();
        int s = 0;
        for (; s + 1 < t; s += 2) {   // two stages per barrier
            // DIF walks the local bits from high to low, DIT from low to high; lbB < lbA are this round's two bits
            const int lbA = DIF ? t - 1 - s : s + 1, lbB = lbA - 1;
            for (int q = threadIdx.x; q < (tile >> 2); q += blockDim.x) {
                const int low = q & ((1 << lbB) - 1);
                const int i00 = ((q >> lbB) << (lbB + 2)) | low, i01 = i00 + (1 << lbB), i10 = i00 + (1 << lbA), i11 = i10 + (1 << lbB);
                Fr x00 = smem_get(plo, phi, i00), x01 = smem_get(plo, phi, i01), x10 = smem_get(plo, phi, i10), x11 = smem_get(plo, phi, i11);
                const int ta = (1 << lbA) - 1, tb = (1 << lbB) - 1;
                if (DIF) {   // bit lbA first: (x00,x10) and (x01,x11), then bit lbB: (x00,x01) and (x10,x11) share a twiddle
                    ntt_bfly<true>(x00, x10, smem_get(wlo, whi, ta + low), false);
                    ntt_bfly<true>(x01, x11, smem_get(wlo, whi, ta + low + (1 << lbB)), false);
                    const bool triv = P.bit_lo + lbB == 0;   // the stage that pairs neighbours: every twiddle is 1
                    const Fr wb = smem_get(wlo, whi, tb + low);
                    ntt_bfly<true>(x00, x01, wb, triv);
                    ntt_bfly<true>(x10, x11, wb, triv);
                } else {     // bit lbB first (shared twiddle), then bit lbA
                    const bool triv = P.bit_lo + lbB == 0;
                    const Fr wb = smem_get(wlo, whi, tb + low);
                    ntt_bfly<false>(x00, x01, wb, triv);
                    ntt_bfly<false>(x10, x11, wb, triv);
                    ntt_bfly<false>(x00, x10, smem_get(wlo, whi, ta + low), false);
                    ntt_bfly<false>(x01, x11, smem_get(wlo, whi, ta + low + (1 << lbB)), false);
                }
                smem_put(plo, phi, i00, x00);
                smem_put(plo, phi, i01, x01);
                smem_put(plo, phi, i10, x10);
                smem_put(plo, phi, i11, x11);
            }
            __syncthreads();
        }
        if (s < t) {   // odd number of stages: one radix-2 stage left
            const int lb = DIF ? 0 : t - 1;
            const bool triv = P.bit_lo + lb == 0;
            for (int bf = threadIdx.x; bf < (tile >> 1); bf += blockDim.x) {
                const int low = bf & ((1 << lb) - 1);
                const int i0 = ((bf >> lb) << (lb + 1)) | low, i1 = i0 + (1 << lb);
                Fr u = smem_get(plo, phi, i0), v = smem_get(plo, phi, i1);
                ntt_bfly<DIF>(u, v, smem_get(wlo, whi, (1 << lb) - 1 + low), triv);
                smem_put(plo, phi, i0, u);
                smem_put(plo, phi, i1, v);
            }
            __syncthreads();
        }
        for (int j = threadIdx.x; j < tile; j += blockDim.x) {
            Fr v = smem_get(plo, phi, j);
            if (P.scale_store) v = fr_mul(v, P.scale);
            fr_store(dst + base + (size_t(j) << P.bit_lo), v);
        }
    }
}

__global__ void __launch_bounds__(256) k_broadcast_rows(const Fr* __restrict__ src, Fr* __restrict__ dst, size_t width,
                                                        size_t out_rows) {
    size_t total = width * out_rows;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x)
        fr_store(dst + i, fr_load(src + i / out_rows));
}

int twiddles(lsp_ctx* ctx, int log_n, bool inverse, const Fr** out) {
    auto& cache = inverse ? ctx->tw_inv : ctx->tw_fwd;
    auto it = cache.find(log_n);
    if (it != cache.end()) {
        *out = it->second;
        return LSP_OK;
    }
    if (log_n < 1 || log_n > 31) return set_err(ctx, LSP_ERR_PARAM, "twiddle size 2^%d unsupported", log_n);
    size_t half = (size_t(1) << log_n) >> 1;
    Fr* tw = nullptr;
    LSP_CUDA(ctx, cudaMalloc(&tw, half * 32));
    LSP_LAUNCH(ctx, k_gen_twiddles, grid_for(ctx, half, 128), 128, 0, (const FieldConsts*)ctx->fc, tw, log_n, inverse ? 1 : 0);
    cache[log_n] = tw;
    *out = tw;
    return LSP_OK;
}

static void split_passes(int log_n, std::vector<int>& ts) {
    int k = (log_n + NTT_MAX_T - 1) / NTT_MAX_T;
    ts.clear();
    for (int i = 0; i < k; i++) ts.push_back(log_n / k + (i < log_n % k ? 1 : 0));
}

// Columns a block loops over: all of a small matrix, at most 8 (the coset factors and twiddles are then amortised 8 x).
constexpr int NTT_COLS_PER_BLOCK = 8;

static int launch_pass(lsp_ctx* ctx, bool dif, NttPass P, size_t width, int n_z) {
    size_t tiles = (size_t(1) << P.log_n) >> P.t;
    P.width = int(width);
    P.n_z = n_z;
    P.cols_per_block = int(width < size_t(NTT_COLS_PER_BLOCK) ? width : size_t(NTT_COLS_PER_BLOCK));
    // a matrix of few columns and few tiles would leave SMs idle: fewer columns per block then
    while (P.cols_per_block > 1 && tiles * size_t(n_z) * ((width + P.cols_per_block - 1) / P.cols_per_block) < size_t(ctx->sm_count) * NTT_TILE_MINB)
        P.cols_per_block = (P.cols_per_block + 1) / 2;
    int threads = 1 << (P.t > 2 ? P.t - 2 : 0);
    if (threads < 32) threads = 32;
    const size_t smem = (size_t(P.pow_lo ? 6 : 4) << P.t) * sizeof(uint4);
    dim3 grid((unsigned)(tiles * size_t(n_z)), (unsigned)((width + P.cols_per_block - 1) / P.cols_per_block), 1);
    if (!ctx->ntt_smem_opt_in) {   // opt in to > 48 KiB of dynamic shared memory, once per device
        LSP_CUDA(ctx, cudaFuncSetAttribute(k_ntt_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (6 << NTT_MAX_T) * int(sizeof(uint4))));
        LSP_CUDA(ctx, cudaFuncSetAttribute(k_ntt_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (6 << NTT_MAX_T) * int(sizeof(uint4))));
        ctx->ntt_smem_opt_in = true;
    }
    if (dif)
        LSP_LAUNCH(ctx, k_ntt_tile<true>, grid, threads, smem, P);
    else
        LSP_LAUNCH(ctx, k_ntt_tile<false>, grid, threads, smem, P);
    return LSP_OK;
}


// Coefficients (natural order, true scale) of every column of `in` (N x W, column-major).
int interpolate_columns(lsp_ctx* ctx, const Fr* in, size_t n, size_t width, Fr* coeffs) {
    int log_n = ilog2(n);
    if (log_n == 0) {
        LSP_CUDA(ctx, cudaMemcpyAsync(coeffs, in, width * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        return LSP_OK;
    }
    const Fr* tw = nullptr;
    LSP_TRY(twiddles(ctx, log_n, true, &tw));
    std::vector<int> ts;
    split_passes(log_n, ts);
    int bit = 0;
    for (size_t p = 0; p < ts.size(); p++) {
        NttPass P;
        memset(&P, 0, sizeof P);
        P.src = p == 0 ? in : coeffs;
        P.dst = coeffs;
        P.src_col_stride = P.dst_col_stride = n;
        P.tw = tw;
        P.log_n = log_n;
        P.bit_lo = bit;
        P.t = ts[p];
        P.bitrev_load = p == 0;
        P.scale_store = p + 1 == ts.size();
        P.scale = host_pow2_inverse(log_n);
        LSP_TRY(launch_pass(ctx, false, P, width, 1));
        bit += ts[p];
    }
    return LSP_OK;
}

// Row blocks [block0, block0 + n_blocks) of the bit-reversed LDE (each block = N rows = one
// coset) from coefficients (N x W).  `out` is column-major with `out_col_stride` rows per
// column and receives the blocks back to back.  block0 = 0, n_blocks = 2^added_bits is the
// whole LDE; a rank of a sharded prove asks for its own range only.
int coset_evaluate_blocks(lsp_ctx* ctx, const Fr* coeffs, size_t n, size_t width, int added_bits, const Fr* shift, int block0,
                          int n_blocks, Fr* out, size_t out_col_stride) {
    int log_n = ilog2(n);
    if (log_n == 0) {  // a constant polynomial
        for (size_t c = 0; c < width; c++)
            LSP_LAUNCH(ctx, k_broadcast_rows, grid_for(ctx, size_t(n_blocks), 256), 256, 0, coeffs + c, out + c * out_col_stride, size_t(1),
                       size_t(n_blocks));
        return LSP_OK;
    }
    const Fr* tw = nullptr;
    LSP_TRY(twiddles(ctx, log_n, false, &tw));
    int lo_bits = log_n < 10 ? log_n : 10;
    size_t n_lo = size_t(1) << lo_bits, n_hi = size_t(1) << (log_n - lo_bits);
    Fr *pow_lo = nullptr, *pow_hi = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&pow_lo, n_blocks * n_lo * 32));
    LSP_TRY(tmp.get((void**)&pow_hi, n_blocks * n_hi * 32));
    {
        dim3 grid((unsigned)((n_lo + n_hi + 127) / 128), (unsigned)n_blocks);
        LSP_LAUNCH(ctx, k_coset_pow_tables, grid, 128, 0, (const FieldConsts*)ctx->fc, pow_lo, pow_hi, shift, log_n, added_bits, lo_bits, block0);
    }
    std::vector<int> ts;
    split_passes(log_n, ts);
    int bit = log_n;
    for (size_t p = 0; p < ts.size(); p++) {
        bit -= ts[p];
        NttPass P;
        memset(&P, 0, sizeof P);
        P.dst = out;
        P.dst_col_stride = out_col_stride;
        P.dst_z_stride = n;
        if (p == 0) {  // coefficients (shared by every coset) -> destination block z, scaled by the coset powers
            P.src = coeffs;
            P.src_col_stride = n;
            P.src_z_stride = 0;
            P.pow_lo = pow_lo;
            P.pow_hi = pow_hi;
            P.pow_lo_bits = lo_bits;
        } else {  // in place on the destination blocks
            P.src = out;
            P.src_col_stride = out_col_stride;
            P.src_z_stride = n;
        }
        P.tw = tw;
        P.log_n = log_n;
        P.bit_lo = bit;
        P.t = ts[p];
        LSP_TRY(launch_pass(ctx, true, P, width, n_blocks));
    }
    return LSP_OK;
}

// ---- a FRACTION of a coset (sharding over more ranks than there are cosets) -------------------------
// Rows [sub*M, (sub+1)*M), M = N >> log_s, of destination block `block`: in bit-reversed order they are the points
// p(S_c * w_N^(k0 + 2^log_s * m)), m < M, with k0 = bitrev_{log_s}(sub) -- the coset sigma * H_M, sigma = S_c * w_N^k0.
// On it x^M = sigma^M =: u is constant, so p folds to degree < M:
//     p(x) = sum_{i0 < M} x^i0 * sum_{t < 2^log_s} a[i0 + M t] u^t
// (the "DIF pre-pass of log2(G/B) stages"), and the rest is the ordinary coset evaluation of the folded polynomial with
// shift sigma.
__global__ void k_subblock_consts(const FieldConsts* __restrict__ fc, const Fr* __restrict__ shift, int log_n, int added_bits, int block,
                                  int log_s, int sub, Fr* __restrict__ out /* [sigma, u] */) {
    const uint32_t coset = bitrev32(uint32_t(block), added_bits), k0 = bitrev32(uint32_t(sub), log_s);
    Fr sigma = fr_mul(fr_load(shift), fr_pow_u32(fr_two_adic_generator(fc, log_n + added_bits), coset));
    sigma = fr_mul(sigma, fr_pow_u32(fr_two_adic_generator(fc, log_n), k0));
    Fr u = sigma;
    for (int i = 0; i < log_n - log_s; i++) u = fr_sqr(u);
    fr_store(out, sigma);
    fr_store(out + 1, u);
}
__global__ void __launch_bounds__(128) k_fold_coeffs(const Fr* __restrict__ coeffs, size_t n, size_t m, int s, const Fr* __restrict__ u_dev,
                                                     Fr* __restrict__ out, size_t width) {
    const Fr u = fr_load(u_dev);
    const size_t total = m * width;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const size_t c = i / m, i0 = i - c * m;
        const Fr* a = coeffs + c * n + i0;
        Fr acc = fr_load_nc(a + size_t(s - 1) * m);
        for (int t = s - 2; t >= 0; t--) acc = fr_add(fr_mul(acc, u), fr_load_nc(a + size_t(t) * m));
        fr_store(out + i, acc);
    }
}
int coset_evaluate_subblock(lsp_ctx* ctx, const Fr* coeffs, size_t n, size_t width, int added_bits, const Fr* shift, int block, int log_s,
                            int sub, Fr* out, size_t out_col_stride) {
    const int log_n = ilog2(n);
    if (log_s < 0 || log_s > log_n) return set_err(ctx, LSP_ERR_PARAM, "a coset of 2^%d rows cannot be split 2^%d ways", log_n, log_s);
    const size_t m = n >> log_s;
    Fr *sc = nullptr, *folded = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&sc, 64));
    LSP_TRY(tmp.get((void**)&folded, m * width * 32));
    LSP_LAUNCH(ctx, k_subblock_consts, 1, 1, 0, (const FieldConsts*)ctx->fc, shift, log_n, added_bits, block, log_s, sub, sc);
    LSP_LAUNCH(ctx, k_fold_coeffs, grid_for(ctx, m * width, 128), 128, 0, coeffs, n, m, 1 << log_s, (const Fr*)(sc + 1), folded, width);
    return coset_evaluate_blocks(ctx, folded, m, width, 0, sc, 0, 1, out, out_col_stride);
}

// out (L x W, column-major, bit-reversed row order) from coefficients (N x W).
int coset_evaluate(lsp_ctx* ctx, const Fr* coeffs, size_t n, size_t width, int added_bits, const Fr* shift, Fr* out) {
    return coset_evaluate_blocks(ctx, coeffs, n, width, added_bits, shift, 0, 1 << added_bits, out, n << added_bits);
}

}  // namespace lsp

extern "C" int lsp_coset_lde_batch(lsp_ctx* ctx, const lsp_mat* in, int added_bits, const uint64_t shift[4],
                                   lsp_mat** out_bitrev, lsp_mat** coeffs_out) {
    if (!ctx || !in || !shift || !out_bitrev || added_bits < 0 || added_bits > 8) return LSP_ERR_PARAM;
    if (!is_pow2(in->rows)) return set_err(ctx, LSP_ERR_PARAM, "matrix height %zu is not a power of two", in->rows);
    if (ilog2(in->rows) + added_bits > 31) return set_err(ctx, LSP_ERR_PARAM, "LDE of 2^%d rows unsupported", ilog2(in->rows) + added_bits);
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = in->rows, w = in->width;
    lsp_mat *co = nullptr, *out = nullptr;
    LSP_TRY(mat_alloc(ctx, n, w, &co));
    int rc = mat_alloc(ctx, n << added_bits, w, &out);
    if (rc == LSP_OK) rc = interpolate_columns(ctx, in->d, n, w, co->d);
    Fr* s_dev = nullptr;
    if (rc == LSP_OK) rc = dev_alloc(ctx, (void**)&s_dev, 32);
    if (rc == LSP_OK && cudaMemcpyAsync(s_dev, shift, 32, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = LSP_ERR_CUDA;
    if (rc == LSP_OK) rc = cudaStreamSynchronize(ctx->stream) == cudaSuccess ? LSP_OK : LSP_ERR_CUDA;  // shift[] is caller-owned
    if (rc == LSP_OK) rc = coset_evaluate(ctx, co->d, n, w, added_bits, s_dev, out->d);
    dev_free(ctx, s_dev);
    if (rc != LSP_OK) {
        lsp_mat_free(ctx, co);
        lsp_mat_free(ctx, out);
        return rc;
    }
    *out_bitrev = out;
    if (coeffs_out)
        *coeffs_out = co;
    else
        lsp_mat_free(ctx, co);
    return LSP_OK;
}
