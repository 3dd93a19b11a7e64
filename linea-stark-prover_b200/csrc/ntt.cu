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
// in shared memory as two 16-byte planes (conflict-free 128-bit accesses).
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

// One block = one tile of 2^t elements of one column (of one coset); blockDim.x = max(2^(t-1), 32).
template <bool DIF>
__global__ void __launch_bounds__(512) k_ntt_pass(const __grid_constant__ NttPass P) {
    extern __shared__ uint4 smem[];
    const int t = P.t;
    const int tile = 1 << t;
    uint4* plo = smem;
    uint4* phi = smem + tile;
    const size_t lo_mask = (size_t(1) << P.bit_lo) - 1;
    const size_t tile_id = blockIdx.x;
    const size_t idx_lo = tile_id & lo_mask;
    const size_t idx_hi = tile_id >> P.bit_lo;
    const size_t base = (idx_hi << (P.bit_lo + t)) | idx_lo;  // + j << bit_lo
    const int z = blockIdx.z;
    const Fr* src = P.src + blockIdx.y * P.src_col_stride + z * P.src_z_stride;
    Fr* dst = P.dst + blockIdx.y * P.dst_col_stride + z * P.dst_z_stride;

    for (int j = threadIdx.x; j < tile; j += blockDim.x) {
        size_t g = base + (size_t(j) << P.bit_lo);
        size_t sidx = P.bitrev_load ? size_t(bitrev32(uint32_t(g), P.log_n)) : g;
        Fr v = fr_load_nc(src + sidx);
        if (P.pow_lo) {
            size_t n_lo = size_t(1) << P.pow_lo_bits, n_hi = size_t(1) << (P.log_n - P.pow_lo_bits);
            Fr a = fr_load_nc(P.pow_lo + z * n_lo + (g & (n_lo - 1)));
            Fr b = fr_load_nc(P.pow_hi + z * n_hi + (g >> P.pow_lo_bits));
            v = fr_mul(v, fr_mul(a, b));
        }
        smem_put(plo, phi, j, v);
    }
    __syncthreads();

    const int nbf = tile >> 1;
    for (int s = 0; s < t; s++) {
        // DIF walks the local bits from high to low, DIT from low to high
        const int lb = DIF ? (t - 1 - s) : s;
        const int b = P.bit_lo + lb;  // global index bit paired at this stage
        for (int bf = threadIdx.x; bf < nbf; bf += blockDim.x) {
            int i0 = ((bf >> lb) << (lb + 1)) | (bf & ((1 << lb) - 1));
            int i1 = i0 + (1 << lb);
            size_t g0 = base + (size_t(i0) << P.bit_lo);
            size_t e = (g0 & ((size_t(1) << b) - 1)) << (P.log_n - 1 - b);
            Fr u = smem_get(plo, phi, i0), v = smem_get(plo, phi, i1);
            if (b == 0) {  // the stage that pairs neighbours: every twiddle is w^0 = 1 (1/log2 n of all products)
                smem_put(plo, phi, i0, fr_add(u, v));
                smem_put(plo, phi, i1, fr_sub(u, v));
                continue;
            }
            Fr w = fr_load_nc(P.tw + e);
            if (DIF) {
                smem_put(plo, phi, i0, fr_add(u, v));
                smem_put(plo, phi, i1, fr_mul(fr_sub(u, v), w));
            } else {
                Fr vw = fr_mul(v, w);
                smem_put(plo, phi, i0, fr_add(u, vw));
                smem_put(plo, phi, i1, fr_sub(u, vw));
            }
        }
        __syncthreads();
    }

    for (int j = threadIdx.x; j < tile; j += blockDim.x) {
        size_t g = base + (size_t(j) << P.bit_lo);
        Fr v = smem_get(plo, phi, j);
        if (P.scale_store) v = fr_mul(v, P.scale);
        fr_store(dst + g, v);
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

static int launch_pass(lsp_ctx* ctx, bool dif, const NttPass& P, size_t width, int n_z) {
    size_t tiles = (size_t(1) << P.log_n) >> P.t;
    int threads = 1 << (P.t > 0 ? P.t - 1 : 0);
    if (threads < 32) threads = 32;
    size_t smem = (size_t(2) << P.t) * sizeof(uint4);
    dim3 grid((unsigned)tiles, (unsigned)width, (unsigned)n_z);
    if (dif)
        LSP_LAUNCH(ctx, k_ntt_pass<true>, grid, threads, smem, P);
    else
        LSP_LAUNCH(ctx, k_ntt_pass<false>, grid, threads, smem, P);
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
// shift sigma.  `next` != 0 evaluates p(w_N * x) instead: the NEXT trace row of every point (the quotient's second operand).
__global__ void k_subblock_consts(const FieldConsts* __restrict__ fc, const Fr* __restrict__ shift, int log_n, int added_bits, int block,
                                  int log_s, int sub, int next, Fr* __restrict__ out /* [sigma, u] */) {
    const uint32_t coset = bitrev32(uint32_t(block), added_bits), k0 = bitrev32(uint32_t(sub), log_s) + (next ? 1u : 0u);
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
                            int sub, bool next, Fr* out, size_t out_col_stride) {
    const int log_n = ilog2(n);
    if (log_s < 0 || log_s > log_n) return set_err(ctx, LSP_ERR_PARAM, "a coset of 2^%d rows cannot be split 2^%d ways", log_n, log_s);
    const size_t m = n >> log_s;
    Fr *sc = nullptr, *folded = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&sc, 64));
    LSP_TRY(tmp.get((void**)&folded, m * width * 32));
    LSP_LAUNCH(ctx, k_subblock_consts, 1, 1, 0, (const FieldConsts*)ctx->fc, shift, log_n, added_bits, block, log_s, sub, next ? 1 : 0, sc);
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
