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
                if (num) fr_store(num + i, fr_mul(fr_load(num + i), inv));
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

// The scans run over a monoid: ADD = false is the running product of the permutation argument
// (trace/src/permutation.rs:72), ADD = true the running log-derivative sum of the lookup (trace/src/lookup.rs:161).
template <bool ADD>
__device__ __forceinline__ Fr scan_id() { return ADD ? fr_zero() : fr_one(); }
template <bool ADD>
__device__ __forceinline__ Fr scan_op(const Fr& a, const Fr& b) { return ADD ? fr_add(a, b) : fr_mul(a, b); }

// Phase 1: in-tile inclusive scan (in place) + tile totals.
template <bool ADD>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(Fr* __restrict__ v, size_t n, Fr* __restrict__ totals) {
    __shared__ uint4 sm[2 * SCAN_THREADS];
    const size_t t0 = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_PER_THREAD;
    Fr x[SCAN_PER_THREAD];
    Fr acc = scan_id<ADD>();
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        x[k] = (t0 + k < n) ? fr_load(v + t0 + k) : scan_id<ADD>();
        acc = scan_op<ADD>(acc, x[k]);
        x[k] = acc;
    }
    sm_put(sm, threadIdx.x, acc);
    __syncthreads();
    for (int off = 1; off < SCAN_THREADS; off <<= 1) {  // Hillis-Steele over the per-thread totals
        Fr mine = sm_get(sm, threadIdx.x), other = scan_id<ADD>();
        bool has = int(threadIdx.x) >= off;
        if (has) other = sm_get(sm, threadIdx.x - off);
        __syncthreads();
        if (has) sm_put(sm, threadIdx.x, scan_op<ADD>(other, mine));
        __syncthreads();
    }
    Fr before = threadIdx.x ? sm_get(sm, threadIdx.x - 1) : scan_id<ADD>();
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++)
        if (t0 + k < n) fr_store(v + t0 + k, threadIdx.x ? scan_op<ADD>(before, x[k]) : x[k]);
    if (threadIdx.x == SCAN_THREADS - 1) fr_store(totals + blockIdx.x, sm_get(sm, SCAN_THREADS - 1));
}
// Phase 2: exclusive scan of the tile totals by one block (sequential over chunks of 256).
template <bool ADD>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_totals(Fr* __restrict__ totals, size_t n_tiles) {
    __shared__ uint4 sm[2 * SCAN_THREADS];
    Fr carry = scan_id<ADD>();
    for (size_t base = 0; base < n_tiles; base += SCAN_THREADS) {
        size_t i = base + threadIdx.x;
        Fr mine = i < n_tiles ? fr_load(totals + i) : scan_id<ADD>();
        sm_put(sm, threadIdx.x, mine);
        __syncthreads();
        for (int off = 1; off < SCAN_THREADS; off <<= 1) {
            Fr cur = sm_get(sm, threadIdx.x), other = scan_id<ADD>();
            bool has = int(threadIdx.x) >= off;
            if (has) other = sm_get(sm, threadIdx.x - off);
            __syncthreads();
            if (has) sm_put(sm, threadIdx.x, scan_op<ADD>(other, cur));
            __syncthreads();
        }
        Fr excl = threadIdx.x ? sm_get(sm, threadIdx.x - 1) : scan_id<ADD>();
        Fr last = sm_get(sm, SCAN_THREADS - 1);
        if (i < n_tiles) fr_store(totals + i, scan_op<ADD>(carry, excl));
        carry = scan_op<ADD>(carry, last);
        __syncthreads();
    }
}
// Phase 3: multiply every tile by the product of the tiles before it.
template <bool ADD>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(Fr* __restrict__ v, size_t n, const Fr* __restrict__ totals) {
    if (blockIdx.x == 0) return;
    Fr pre = fr_load(totals + blockIdx.x);
    const size_t t0 = size_t(blockIdx.x) * SCAN_TILE;
    for (int k = threadIdx.x; k < SCAN_TILE; k += SCAN_THREADS)
        if (t0 + k < n) fr_store(v + t0 + k, scan_op<ADD>(pre, fr_load(v + t0 + k)));
}

__global__ void k_check_last_is_one(const Fr* __restrict__ v, size_t n, int* __restrict__ flag) {
    *flag = fr_eq(fr_load(v + n - 1), fr_one()) ? 0 : 1;
}
__global__ void k_check_last_is_zero(const Fr* __restrict__ v, size_t n, int* __restrict__ flag) {
    *flag = fr_is_zero(fr_load(v + n - 1)) ? 0 : 1;
}

// ==========================================================================================
// LogUp lookup witness: `RawLookupTrace::get_trace` (reference trace/src/lookup.rs:46-176).
//   columns  a.., b (table by table).., a_filter, b_filter[T], 1/(a_comb+delta), 1/(b_comb+delta)[T],
//            multiplicities[T], running sum of  filter_a/(a_comb+delta) - sum_t mult_t/(b_comb_t+delta)
// The reference counts the enabled A rows in a HashMap keyed by the Horner combination (:80-104)
// and hands each count to the FIRST enabled B row (row-major scan, tables in order) holding that
// key, removing it afterwards (:139-154).  Here: an open-addressing table whose slots store a
// representative ROW INDEX (the 256-bit keys already sit in an array, so claiming a slot is one
// 32-bit CAS and equality is a full-width compare: exact, lock-free, no waiting); counts by
// atomicAdd; the first enabled B row per key by atomicMin over i*T+t.  Order-independent, hence
// deterministic.
// ==========================================================================================
struct LookupWitArgs {
    const Fr* in_rm;     // host layout, row-major rows x w_in, w_in = n_a + T*n_b + 1 + T
    size_t n;
    int n_a, n_t, n_b;
    const Fr* publics;   // alpha, delta
    Fr* out;             // column-major n x (n_a + T*(n_b+3) + 3)
    Fr* a_key;           // n           a_comb (no delta)
    Fr* b_key;           // T x n       b_comb
    int* rep;            // cap         slot -> representative A row (-1 empty)
    unsigned* count;     // cap
    unsigned long long* first;  // cap  min i*T+t over enabled matching B rows
    unsigned cap_mask;
};

__device__ __forceinline__ unsigned lk_hash(const Fr& k) {
    unsigned h = k.l[0] * 0x9e3779b1u;
    h ^= (k.l[1] + 0x7f4a7c15u) * 0x85ebca6bu;
    h ^= (k.l[3] >> 3) * 0xc2b2ae35u;
    h ^= k.l[6] * 0x27d4eb2fu;
    return h ^ (h >> 15);
}

__global__ void __launch_bounds__(128) k_lk_combine(const __grid_constant__ LookupWitArgs A) {
    const Fr alpha = fr_load(A.publics), delta = fr_load(A.publics + 1);
    const int w_in = A.n_a + A.n_t * A.n_b + 1 + A.n_t;
    const int col_a_inv = w_in;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < A.n; i += size_t(gridDim.x) * blockDim.x) {
        const Fr* row = A.in_rm + i * size_t(w_in);
        Fr acc = fr_zero();
        for (int j = 0; j < A.n_a; j++) {                       // :122-127
            Fr v = fr_load_nc(row + j);
            fr_store(A.out + size_t(j) * A.n + i, v);
            acc = j ? fr_add(fr_mul(acc, alpha), v) : v;
        }
        fr_store(A.a_key + i, acc);
        fr_store(A.out + size_t(col_a_inv) * A.n + i, fr_add(acc, delta));            // inverted later (:129)
        for (int t = 0; t < A.n_t; t++) {
            Fr bc = fr_zero();
            for (int j = 0; j < A.n_b; j++) {                   // :139-144
                const int c = A.n_a + t * A.n_b + j;
                Fr v = fr_load_nc(row + c);
                fr_store(A.out + size_t(c) * A.n + i, v);
                bc = j ? fr_add(fr_mul(bc, alpha), v) : v;
            }
            fr_store(A.b_key + size_t(t) * A.n + i, bc);
            fr_store(A.out + size_t(col_a_inv + 1 + t) * A.n + i, fr_add(bc, delta));  // :146
        }
        for (int f = 0; f < 1 + A.n_t; f++) {                   // filters (:69-70)
            const int c = A.n_a + A.n_t * A.n_b + f;
            fr_store(A.out + size_t(c) * A.n + i, fr_load_nc(row + c));
        }
    }
}

// slot of `key`, inserting row `ins` as its representative when absent (ins < 0: lookup only; -1 when absent)
__device__ __forceinline__ int lk_slot(const LookupWitArgs& A, const Fr& key, int ins) {
    unsigned s = lk_hash(key) & A.cap_mask;
    for (unsigned probe = 0; probe <= A.cap_mask; probe++, s = (s + 1) & A.cap_mask) {
        int cur = ins >= 0 ? atomicCAS(A.rep + s, -1, ins) : A.rep[s];
        if (cur == -1) return ins >= 0 ? int(s) : -1;
        if (fr_eq(fr_load(A.a_key + cur), key)) return int(s);
    }
    return -1;
}

__global__ void __launch_bounds__(128) k_lk_insert(const __grid_constant__ LookupWitArgs A) {
    const Fr* a_filter = A.out + size_t(A.n_a + A.n_t * A.n_b) * A.n;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < A.n; i += size_t(gridDim.x) * blockDim.x) {
        if (fr_is_zero(fr_load(a_filter + i))) continue;        // :85-87
        int s = lk_slot(A, fr_load(A.a_key + i), int(i));
        if (s >= 0) atomicAdd(A.count + s, 1u);                 // :98-103
    }
}

// pass 0: the first enabled B row of every key;  pass 1: multiplicities
template <int PASS>
__global__ void __launch_bounds__(128) k_lk_match(const __grid_constant__ LookupWitArgs A) {
    const int w_in = A.n_a + A.n_t * A.n_b + 1 + A.n_t;
    const size_t total = A.n * size_t(A.n_t);
    for (size_t e = blockIdx.x * size_t(blockDim.x) + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
        const size_t i = e / A.n_t;
        const int t = int(e - i * A.n_t);
        const Fr* b_filter = A.out + size_t(A.n_a + A.n_t * A.n_b + 1 + t) * A.n;
        const bool enabled = !fr_is_zero(fr_load(b_filter + i));
        int s = enabled ? lk_slot(A, fr_load(A.b_key + size_t(t) * A.n + i), -1) : -1;
        if (PASS == 0) {
            if (s >= 0) atomicMin(A.first + s, (unsigned long long)e);
        } else {
            Fr occ = fr_zero();                                 // :148
            if (s >= 0 && A.first[s] == (unsigned long long)e) {  // :149-156: count goes to the first enabled holder
                occ.l[0] = A.count[s];
                occ = fr_mul(occ, fr_const(FR_R2));             // from_canonical_usize
            }
            fr_store(A.out + size_t(w_in + 1 + A.n_t + t) * A.n + i, occ);
        }
    }
}

// term_i = [a_filter_i != 0] * a_inv_i - sum_t mult_{t,i} * b_inv_{t,i}    (:133-136,152)
__global__ void __launch_bounds__(128) k_lk_terms(const __grid_constant__ LookupWitArgs A) {
    const int w_in = A.n_a + A.n_t * A.n_b + 1 + A.n_t;
    const Fr* a_filter = A.out + size_t(A.n_a + A.n_t * A.n_b) * A.n;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < A.n; i += size_t(gridDim.x) * blockDim.x) {
        Fr term = fr_is_zero(fr_load(a_filter + i)) ? fr_zero() : fr_load(A.out + size_t(w_in) * A.n + i);
        for (int t = 0; t < A.n_t; t++) {
            Fr m = fr_load(A.out + size_t(w_in + 1 + A.n_t + t) * A.n + i);
            if (!fr_is_zero(m)) term = fr_sub(term, fr_mul(m, fr_load(A.out + size_t(w_in + 1 + t) * A.n + i)));
        }
        fr_store(A.out + size_t(w_in + 1 + 2 * A.n_t) * A.n + i, term);
    }
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
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&stage, n * w_in * 32));
    LSP_TRY(tmp.get((void**)&pub, 64));
    LSP_TRY(tmp.get((void**)&totals, n_tiles * 32));
    LSP_TRY(tmp.get((void**)&flag, 4));
    lsp_mat* m = nullptr;
    LSP_TRY(mat_alloc(ctx, n, w_out, &m));
    MatGuard m_guard{ctx, m};   // released unless handed to the caller
    LSP_CUDA(ctx, cudaMemcpyAsync(stage, ab_rowmajor, n * w_in * 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemcpyAsync(pub, publics, 64, cudaMemcpyHostToDevice, ctx->stream));
    if (big_endian_bytes) LSP_LAUNCH(ctx, k_be_to_mont, grid_for(ctx, n * w_in, 128), 128, 0, stage, n * w_in);
    Fr* den = m->d + w_in * n;
    Fr* num = den + n;
    LSP_LAUNCH(ctx, k_wit_combine, grid_for(ctx, n, 128), 128, 0, (const Fr*)stage, n, int(n_cols), (const Fr*)pub, m->d);
    LSP_LAUNCH(ctx, k_wit_invert, grid_for(ctx, (n + WIT_BATCH - 1) / WIT_BATCH, 128), 128, 0, den, num, n);
    LSP_LAUNCH(ctx, k_scan_tiles<false>, unsigned(n_tiles), SCAN_THREADS, 0, num, n, totals);
    LSP_LAUNCH(ctx, k_scan_totals<false>, 1, SCAN_THREADS, 0, totals, n_tiles);
    LSP_LAUNCH(ctx, k_scan_apply<false>, unsigned(n_tiles), SCAN_THREADS, 0, num, n, (const Fr*)totals);
    LSP_LAUNCH(ctx, k_check_last_is_one, 1, 1, 0, (const Fr*)num, n, flag);
    LSP_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*(volatile int*)ctx->pinned)
        return set_err(ctx, LSP_ERR_PARAM, "failed to check constrain: check column should be 1 on the last row");
    *trace_out = m_guard.release();
    return LSP_OK;
}

static int lookup_trace_impl(lsp_ctx* ctx, const void* in_rowmajor, bool big_endian_bytes, size_t rows, uint32_t n_a, uint32_t n_t,
                             uint32_t n_b, const uint64_t publics[2][4], lsp_mat** trace_out) {
    if (!ctx || !in_rowmajor || !publics || !trace_out || rows == 0 || n_a == 0 || n_t == 0 || n_b == 0) return LSP_ERR_PARAM;
    if (rows > (size_t(1) << 30)) return set_err(ctx, LSP_ERR_PARAM, "lookup trace of %zu rows unsupported", rows);
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = rows, w_in = size_t(n_a) + size_t(n_t) * n_b + 1 + n_t, w_out = w_in + 2 * size_t(n_t) + 2;
    size_t cap = 64;
    while (cap < 2 * n) cap <<= 1;
    const size_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    Fr *stage = nullptr, *pub = nullptr, *totals = nullptr, *a_key = nullptr, *b_key = nullptr;
    int *flag = nullptr, *rep = nullptr;
    unsigned* count = nullptr;
    unsigned long long* first = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&stage, n * w_in * 32));
    LSP_TRY(tmp.get((void**)&pub, 64));
    LSP_TRY(tmp.get((void**)&totals, n_tiles * 32));
    LSP_TRY(tmp.get((void**)&a_key, n * 32));
    LSP_TRY(tmp.get((void**)&b_key, n * n_t * 32));
    LSP_TRY(tmp.get((void**)&flag, 4));
    LSP_TRY(tmp.get((void**)&rep, cap * 4));
    LSP_TRY(tmp.get((void**)&count, cap * 4));
    LSP_TRY(tmp.get((void**)&first, cap * 8));
    lsp_mat* m = nullptr;
    LSP_TRY(mat_alloc(ctx, n, w_out, &m));
    MatGuard m_guard{ctx, m};   // released unless handed to the caller
    LSP_CUDA(ctx, cudaMemcpyAsync(stage, in_rowmajor, n * w_in * 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemcpyAsync(pub, publics, 64, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemsetAsync(rep, 0xff, cap * 4, ctx->stream));
    LSP_CUDA(ctx, cudaMemsetAsync(count, 0, cap * 4, ctx->stream));
    LSP_CUDA(ctx, cudaMemsetAsync(first, 0xff, cap * 8, ctx->stream));
    if (big_endian_bytes) LSP_LAUNCH(ctx, k_be_to_mont, grid_for(ctx, n * w_in, 128), 128, 0, stage, n * w_in);
    LookupWitArgs A;
    A.in_rm = stage;
    A.n = n;
    A.n_a = int(n_a);
    A.n_t = int(n_t);
    A.n_b = int(n_b);
    A.publics = pub;
    A.out = m->d;
    A.a_key = a_key;
    A.b_key = b_key;
    A.rep = rep;
    A.count = count;
    A.first = first;
    A.cap_mask = unsigned(cap - 1);
    Fr* dens = m->d + w_in * n;                 // a_inv and the T b_inv columns are contiguous
    Fr* prefix = m->d + (w_out - 1) * n;
    LSP_LAUNCH(ctx, k_lk_combine, grid_for(ctx, n, 128), 128, 0, A);
    LSP_LAUNCH(ctx, k_lk_insert, grid_for(ctx, n, 128), 128, 0, A);
    LSP_LAUNCH(ctx, k_lk_match<0>, grid_for(ctx, n * n_t, 128), 128, 0, A);
    LSP_LAUNCH(ctx, k_lk_match<1>, grid_for(ctx, n * n_t, 128), 128, 0, A);
    const size_t n_den = n * (1 + size_t(n_t));
    LSP_LAUNCH(ctx, k_wit_invert, grid_for(ctx, (n_den + WIT_BATCH - 1) / WIT_BATCH, 128), 128, 0, dens, (Fr*)nullptr, n_den);
    LSP_LAUNCH(ctx, k_lk_terms, grid_for(ctx, n, 128), 128, 0, A);
    LSP_LAUNCH(ctx, k_scan_tiles<true>, unsigned(n_tiles), SCAN_THREADS, 0, prefix, n, totals);
    LSP_LAUNCH(ctx, k_scan_totals<true>, 1, SCAN_THREADS, 0, totals, n_tiles);
    LSP_LAUNCH(ctx, k_scan_apply<true>, unsigned(n_tiles), SCAN_THREADS, 0, prefix, n, (const Fr*)totals);
    LSP_LAUNCH(ctx, k_check_last_is_zero, 1, 1, 0, (const Fr*)prefix, n, flag);
    LSP_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*(volatile int*)ctx->pinned)
        return set_err(ctx, LSP_ERR_PARAM, "failed to check constrain: check column should be 0 on the last row");
    *trace_out = m_guard.release();
    return LSP_OK;
}

extern "C" int lsp_lookup_trace(lsp_ctx* ctx, const uint64_t* in_rowmajor, size_t rows, uint32_t n_a_cols, uint32_t n_tables,
                                uint32_t n_b_cols, const uint64_t publics[2][4], lsp_mat** trace_out) {
    return lookup_trace_impl(ctx, in_rowmajor, false, rows, n_a_cols, n_tables, n_b_cols, publics, trace_out);
}

extern "C" int lsp_lookup_trace_be(lsp_ctx* ctx, const uint8_t* be_rowmajor, size_t rows, uint32_t n_a_cols, uint32_t n_tables,
                                   uint32_t n_b_cols, const uint64_t publics[2][4], lsp_mat** trace_out) {
    return lookup_trace_impl(ctx, be_rowmajor, true, rows, n_a_cols, n_tables, n_b_cols, publics, trace_out);
}

// `RawTrace::push_traces` column concatenation (trace/src/lib.rs:50-60,80-89): matrices of one height side by side.
extern "C" int lsp_mat_hconcat(lsp_ctx* ctx, const lsp_mat* const* mats, int n_mats, lsp_mat** out) {
    if (!ctx || !mats || n_mats <= 0 || !out) return LSP_ERR_PARAM;
    size_t w = 0;
    for (int i = 0; i < n_mats; i++) {
        if (!mats[i] || mats[i]->rows != mats[0]->rows) return set_err(ctx, LSP_ERR_PARAM, "hconcat: heights differ");
        w += mats[i]->width;
    }
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    lsp_mat* m = nullptr;
    LSP_TRY(mat_alloc(ctx, mats[0]->rows, w, &m));
    size_t off = 0;
    for (int i = 0; i < n_mats; i++) {   // column-major: each matrix is one contiguous block of columns
        const size_t bytes = mats[i]->rows * mats[i]->width * 32;
        LSP_CUDA(ctx, cudaMemcpyAsync(m->d + off, mats[i]->d, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        off += mats[i]->rows * mats[i]->width;
    }
    *out = m;
    return LSP_OK;
}
