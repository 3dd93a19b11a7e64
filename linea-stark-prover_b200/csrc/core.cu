// Context, matrices, field/permutation parity probes and the Merkle MMCS.
#include "stark.cuh"

using namespace lsp;

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_fr_op(int op, const Fr* __restrict__ a, const Fr* __restrict__ b,
                                               Fr* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        Fr x = fr_load(a + i), r;
        switch (op) {
            case 0: r = fr_add(x, fr_load(b + i)); break;
            case 1: r = fr_sub(x, fr_load(b + i)); break;
            case 2: r = fr_mul(x, fr_load(b + i)); break;
            case 3: r = fr_inv(x); break;
            default: r = fr_halve(x); break;
        }
        fr_store(out + i, r);
    }
}

// host row-major (rows x width)  <->  device column-major
__global__ void __launch_bounds__(256) k_rm_to_cm(const Fr* __restrict__ rm, Fr* __restrict__ cm, size_t rows,
                                                  size_t width, size_t row0, size_t nrows) {
    size_t total = nrows * width;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        size_t r = i / width, c = i - r * width;
        fr_store(cm + c * rows + row0 + r, fr_load_nc(rm + i));
    }
}
__global__ void __launch_bounds__(256) k_cm_to_rm(const Fr* __restrict__ cm, Fr* __restrict__ rm, size_t rows,
                                                  size_t width, size_t row0, size_t nrows) {
    size_t total = nrows * width;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        size_t r = i / width, c = i - r * width;
        fr_store(rm + i, fr_load_nc(cm + c * rows + row0 + r));
    }
}

template <int D>
__global__ void __launch_bounds__(128) k_p2_permute(const __grid_constant__ P2Params P, const Fr* __restrict__ in,
                                                    Fr* __restrict__ out, size_t n) {
    LSP_P2_SLOT_DECL(128);
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        Fr s0 = fr_load(in + 3 * i), s1 = fr_load(in + 3 * i + 1), s2 = fr_load(in + 3 * i + 2);
        p2_permute<D, 128>(P, s0, s1, s2, LSP_P2_SLOT(128));
        fr_store(out + 3 * i, s0);
        fr_store(out + 3 * i + 1, s1);
        fr_store(out + 3 * i + 2, s2);
    }
}

// Resident 128-thread blocks per SM the one-thread-per-permutation kernels are compiled for (64 registers, no spills, at 8; measured 4..8: 102.5, 102.0, 100.4, 100.0, 98.7 ms per prove).
#ifndef LSP_P2_MINB
#define LSP_P2_MINB 8
#endif

// Leaf digests: one thread per row.  PaddingFreeSponge<Perm,3,2,1>::hash_iter over
// the concatenation of that row in every matrix (overwrite mode, rate 2).
// cols[] are column base pointers (column-major storage => coalesced across rows).
// The permutation is inlined at ONE site (odd tail folded into the loop): the kernel stays
// inside the instruction cache.
template <int D>
__global__ void __launch_bounds__(128, LSP_P2_MINB) k_leaf_hash(const __grid_constant__ P2Params P, const Fr* const* __restrict__ cols,
                                                   int width, size_t rows, Fr* __restrict__ digests) {
    LSP_P2_SLOT_DECL(128);
    for (size_t r = blockIdx.x * size_t(blockDim.x) + threadIdx.x; r < rows; r += size_t(gridDim.x) * blockDim.x) {
        Fr s0 = fr_zero(), s1 = fr_zero(), s2 = fr_zero();
#pragma unroll 1
        for (int c = 0; c < width; c += 2) {
            s0 = fr_load_nc(cols[c] + r);
            if (c + 1 < width) s1 = fr_load_nc(cols[c + 1] + r);  // odd tail: state[1] keeps its stale value
            p2_permute<D, 128>(P, s0, s1, s2, LSP_P2_SLOT(128));
        }
        fr_store(digests + r, s0);
    }
}

// One Merkle layer: out[i] = compress(in[2i], in[2i+1]).
template <int D>
__global__ void __launch_bounds__(128, LSP_P2_MINB) k_compress_layer(const __grid_constant__ P2Params P, const Fr* __restrict__ in,
                                                        Fr* __restrict__ out, size_t n_out) {
    LSP_P2_SLOT_DECL(128);
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n_out; i += size_t(gridDim.x) * blockDim.x) {
        Fr l = fr_load(in + 2 * i), r = fr_load(in + 2 * i + 1);
        fr_store(out + i, p2_compress<D, 128>(P, l, r, LSP_P2_SLOT(128)));
    }
}

// The same layer for the small tops of the trees, where the cost is the latency of one
// permutation: three lanes per node (p2_permute_tri), one warp per block so that the warps
// spread over the SM sub-partitions.  Lane 3k+w of a warp holds word w of node warp*10+k.
template <int D>
__global__ void __launch_bounds__(32) k_compress_layer_tri(const __grid_constant__ P2Params P, const Fr* __restrict__ in,
                                                           Fr* __restrict__ out, size_t n_out) {
    const int lane = threadIdx.x, k = lane / 3, w = lane - 3 * k;
    size_t i = size_t(blockIdx.x) * 10 + k;
    const bool live = k < 10 && i < n_out;
    if (!live) i = 0;
    Fr s = w < 2 ? fr_load(in + 2 * i + w) : fr_zero();
    p2_permute_tri<D>(P, s, w, 3 * k);
    if (live && w == 0) fr_store(out + i, s);
}

// The whole top of a tree in ONE launch: levels of <= 10 * gridDim.x nodes, three lanes per node, one warp per
// block, a grid-wide barrier between levels (cooperative launch: every block is resident).  Level j reads
// `in` and writes `out`; level j+1 reads that and writes right behind it (digest layers are stored back to
// back).  Saves the ~7 us a separate launch costs per level on a chain of ~280 latency-bound levels per prove.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
    __syncwarp();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*(volatile unsigned*)counter < target) {
        }
        __threadfence();
    }
    __syncwarp();
}
template <int D>
__global__ void __launch_bounds__(32) k_compress_top_tri(const __grid_constant__ P2Params P, const Fr* in, Fr* out, size_t n_out,
                                                         int n_levels, unsigned* barrier) {
    const int lane = threadIdx.x, k = lane / 3, w = lane - 3 * k;
    for (int lvl = 0; lvl < n_levels; lvl++) {
        size_t i = size_t(blockIdx.x) * 10 + k;
        const bool live = k < 10 && i < n_out;
        if (size_t(blockIdx.x) * 10 < n_out) {   // warp-uniform
            if (!live) i = 0;
            Fr s = fr_zero();
            if (w < 2) {   // written by other blocks one level earlier: read through L2
                const uint4* q = reinterpret_cast<const uint4*>(in + 2 * i + w);
                uint4 a = __ldcg(q), b = __ldcg(q + 1);
                s.l[0] = a.x; s.l[1] = a.y; s.l[2] = a.z; s.l[3] = a.w;
                s.l[4] = b.x; s.l[5] = b.y; s.l[6] = b.z; s.l[7] = b.w;
            }
            p2_permute_tri<D>(P, s, w, 3 * k);
            if (live && w == 0) fr_store(out + i, s);
        }
        if (lvl + 1 < n_levels) grid_barrier(barrier, unsigned(lvl + 1) * gridDim.x);
        in = out;
        out += n_out;
        n_out >>= 1;
    }
}

// Gather one row across columns (open_batch) into a contiguous buffer.
__global__ void k_gather_row(const Fr* const* __restrict__ cols, int width, size_t row, Fr* __restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < width) fr_store(out + c, fr_load(cols[c] + row));
}
__global__ void k_gather_siblings(const Fr* __restrict__ digests, size_t h, int log_h, size_t index, Fr* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < log_h) {
        size_t off = 2 * h - ((2 * h) >> k);
        fr_store(out + k, fr_load(digests + off + ((index >> k) ^ 1)));
    }
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
static bool limbs_reduced(const uint64_t* l) {
    static const uint64_t P[4] = {0x0a11800000000001ull, 0x59aa76fed0000001ull, 0x60b44d1e5c37b001ull, 0x12ab655e9a2ca556ull};
    for (int i = 3; i >= 0; i--) {
        if (l[i] < P[i]) return true;
        if (l[i] > P[i]) return false;
    }
    return false;
}

extern "C" int lsp_abi_version(void) { return LSP_ABI_VERSION; }

// Validates and completes the field parameters: gen != 0 with gen^(2^31) != 1 (its cosets of every supported two-adic
// subgroup are disjoint from the subgroup: L <= 2^31), root^(2^46) == -1 (a primitive 2^47-th root of unity), and
// fills gen_inv.  status: 0 ok, 1 bad generator, 2 bad root.
__global__ void k_field_consts_setup(FieldConsts* fc, int* status) {
    const Fr g = fr_load(&fc->gen), w = fr_load(&fc->root47);
    int st = 0;
    Fr t = g;
    for (int i = 0; i < 31; i++) t = fr_sqr(t);
    if (fr_is_zero(g) || fr_eq(t, fr_one())) st = 1;
    t = w;
    for (int i = 0; i < 46; i++) t = fr_sqr(t);
    if (!fr_eq(t, fr_neg(fr_one()))) st = st ? st : 2;
    fr_store(&fc->gen_inv, fr_inv(g));
    *status = st;
}

static void drop_shape_caches(lsp_ctx* ctx) {  // everything derived from the field parameters
    for (auto& kv : ctx->tw_fwd) cudaFree(kv.second);
    for (auto& kv : ctx->tw_inv) cudaFree(kv.second);
    ctx->tw_fwd.clear();
    ctx->tw_inv.clear();
    for (auto& kv : ctx->quot_sel) {
        cudaFree(kv.second.scal);
        cudaFree(kv.second.inv0);
        cudaFree(kv.second.inv1);
    }
    ctx->quot_sel.clear();
}

// `Val::GENERATOR` and `Val::two_adic_generator(47)` (fork-only `p3-bls12-377-fr`), Montgomery limbs.
extern "C" int lsp_set_field_consts(lsp_ctx* ctx, const uint64_t generator[4], const uint64_t two_adic_root_2_47[4]) {
    if (!ctx || !generator || !two_adic_root_2_47) return LSP_ERR_PARAM;
    if (!limbs_reduced(generator) || !limbs_reduced(two_adic_root_2_47)) return set_err(ctx, LSP_ERR_PARAM, "field constant is not reduced");
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FieldConsts old;
    LSP_CUDA(ctx, cudaMemcpy(&old, ctx->fc, sizeof old, cudaMemcpyDeviceToHost));
    FieldConsts h;
    memset(&h, 0, sizeof h);
    memcpy(&h.gen, generator, 32);
    memcpy(&h.root47, two_adic_root_2_47, 32);
    int* status = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&status, 4));
    LSP_CUDA(ctx, cudaMemcpyAsync(ctx->fc, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
    LSP_LAUNCH(ctx, k_field_consts_setup, 1, 1, 0, ctx->fc, status);
    LSP_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, status, 4, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int st = *(volatile int*)ctx->pinned;
    if (st != 0) {
        LSP_CUDA(ctx, cudaMemcpy(ctx->fc, &old, sizeof old, cudaMemcpyHostToDevice));
        return set_err(ctx, LSP_ERR_PARAM, st == 1 ? "generator is zero or lies in a two-adic subgroup" : "not a primitive 2^47-th root of unity");
    }
    drop_shape_caches(ctx);  // twiddles and selector tables were built from the old parameters
    return LSP_OK;
}

// Transcript order of `TwoAdicFriPcs::open`.  The pinned fork: (1, 0).  Later upstream: (0, 1).
extern "C" int lsp_set_transcript_flags(lsp_ctx* ctx, int alpha_before_openings, int observe_opened_values) {
    if (!ctx) return LSP_ERR_PARAM;
    ctx->alpha_before_openings = alpha_before_openings != 0;
    ctx->observe_opened_values = observe_opened_values != 0;
    return LSP_OK;
}

extern "C" int lsp_ctx_create(int device, lsp_ctx** out) {
    if (!out) return LSP_ERR_PARAM;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return LSP_ERR_CUDA;
    lsp_ctx* ctx = new lsp_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return LSP_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    // keep freed blocks in the stream-ordered pool: prove() allocates the same shapes every call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    ctx->pinned_bytes = 1 << 20;
    if (cudaMallocHost(&ctx->pinned, ctx->pinned_bytes) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return LSP_ERR_NOMEM;
    }
    // field parameters: arkworks' FrConfig values until lsp_set_field_consts says otherwise
    int rc = cudaMalloc((void**)&ctx->fc, sizeof(FieldConsts)) == cudaSuccess ? LSP_OK : LSP_ERR_NOMEM;
    if (rc == LSP_OK) {
        uint64_t g[4], w[4];
        memcpy(g, FR_GEN_DEFAULT, 32);
        memcpy(w, FR_ROOT47_DEFAULT, 32);
        cudaMemset(ctx->fc, 0, sizeof(FieldConsts));
        rc = lsp_set_field_consts(ctx, g, w);
    }
    if (rc != LSP_OK) {
        lsp_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return LSP_OK;
}

extern "C" void lsp_ctx_destroy(lsp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    drop_shape_caches(ctx);
    cudaFree(ctx->fc);
    for (auto& r : ctx->timing_recs) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    for (auto& e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaFree(ctx->grid_barrier);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* lsp_last_error(const lsp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

extern "C" int lsp_ctx_sync(lsp_ctx* ctx) {
    if (!ctx) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSP_OK;
}

extern "C" uint64_t lsp_kernel_launches(const lsp_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int lsp_kernel_timing(lsp_ctx* ctx, int enable) {
    if (!ctx) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& r : ctx->timing_recs) {
        ctx->ev_pool.push_back(r.e0);
        ctx->ev_pool.push_back(r.e1);
    }
    ctx->timing_recs.clear();
    ctx->timing = enable != 0;
    ctx->timing_leaf_only = enable == 2;
    return LSP_OK;
}

// JSON: [{"phase": "...", "kernel": "...", "launches": n, "ms": total}, ...] aggregated by (phase, kernel)
extern "C" int lsp_kernel_timing_report(lsp_ctx* ctx, char* buf, size_t cap) {
    if (!ctx || !buf || cap < 3) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<std::pair<std::string, std::pair<int, double>>> agg;
    for (auto& r : ctx->timing_recs) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        std::string key = std::string(r.phase) + "|" + r.name;
        bool found = false;
        for (auto& a : agg)
            if (a.first == key) {
                a.second.first++;
                a.second.second += ms;
                found = true;
                break;
            }
        if (!found) agg.push_back({key, {1, double(ms)}});
    }
    std::string out = "[";
    for (size_t i = 0; i < agg.size(); i++) {
        size_t bar = agg[i].first.find('|');
        char line[512];
        snprintf(line, sizeof line, "%s{\"phase\": \"%s\", \"kernel\": \"%s\", \"launches\": %d, \"ms\": %.6f}", i ? ", " : "",
                 agg[i].first.substr(0, bar).c_str(), agg[i].first.substr(bar + 1).c_str(), agg[i].second.first, agg[i].second.second);
        out += line;
    }
    out += "]";
    if (out.size() + 1 > cap) return set_err(ctx, LSP_ERR_PARAM, "timing report needs %zu bytes", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return LSP_OK;
}

// Integer-pipe peaks of this device: the 32x32->64 multiply-accumulate in each instruction form the
// compiler can emit for it, eight independent accumulators per thread, DATA-DEPENDENT multiplicands
// (with loop-invariant operands ptxas hoists the product and the "multiply" becomes an add).  The SASS
// of each form is committed under profiles/sass_k_int_peak.txt.  Timed with CUDA events on the ctx stream.
//   form 0  IMAD.WIDE.U32 Rd, Ra, Rb, RZ + IADD3 + IADD3.X   multiply only, the 64-bit accumulation on the ALU pipe
//           (what ptxas makes of mad.wide.u32 here: it splits the add off to shorten the multiplicand's dependency)
//   form 1  IMAD.WIDE.U32 Rd, Ra, Rb, Rc                     fused 64-bit multiply-accumulate, no carry in or out
//   form 2  IMAD.WIDE.U32.X Rd, P, Ra, Rb, Rc, P             carry in and out: the link the Montgomery product is made of
//   form 3  IMAD + IMAD.HI.U32                               the pair ptxas falls back to when it cannot fuse
template <int FORM>
__global__ void __launch_bounds__(256) k_int_peak(uint32_t* out, uint32_t seed, int iters) {
    uint32_t b = blockIdx.x * 40503u + 17u + seed + threadIdx.x * 2654435761u;
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        lo[c] = (b + c) * 0x9e3779b9u;
        hi[c] = (b ^ c) * 0x7f4a7c15u;
    }
    for (int it = 0; it < iters; it++) {
        if (FORM == 2) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(b) : "r"(lo[0]));  // opens the carry chain
#pragma unroll
        for (int c = 0; c < 8; c++) {  // (lo,hi)[c] += lo[c+1] * b
            if (FORM == 0) {
                unsigned long long t = (static_cast<unsigned long long>(hi[c]) << 32) | lo[c];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(lo[(c + 1) & 7]), "r"(b));
                lo[c] = uint32_t(t);
                hi[c] = uint32_t(t >> 32);
            } else if (FORM == 1)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(lo[(c + 1) & 7]), "r"(b));
            else if (FORM == 2)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(lo[(c + 1) & 7]), "r"(b));
            else
                asm volatile("mad.lo.u32 %0, %2, %3, %0;\n\tmad.hi.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(hi[(c + 1) & 7]), "r"(b));
        }
    }
    uint32_t r = b;
#pragma unroll
    for (int c = 0; c < 8; c++) r ^= lo[c] ^ hi[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

extern "C" int lsp_int_peaks(lsp_ctx* ctx, double mac32_per_s[LSP_INT_PEAK_FORMS]) {
    if (!ctx || !mac32_per_s) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 8192;
    uint32_t* out = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&out, size_t(blocks) * threads * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int form = 0; form < LSP_INT_PEAK_FORMS; form++) {
        float best = 1e30f;
        for (int r = 0; r < 6; r++) {
            cudaEventRecord(e0, ctx->stream);
            switch (form) {
                case 0: k_int_peak<0><<<blocks, threads, 0, ctx->stream>>>(out, uint32_t(r), iters); break;
                case 1: k_int_peak<1><<<blocks, threads, 0, ctx->stream>>>(out, uint32_t(r), iters); break;
                case 2: k_int_peak<2><<<blocks, threads, 0, ctx->stream>>>(out, uint32_t(r), iters); break;
                default: k_int_peak<3><<<blocks, threads, 0, ctx->stream>>>(out, uint32_t(r), iters); break;
            }
            cudaEventRecord(e1, ctx->stream);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 2 && ms < best) best = ms;
        }
        ctx->launches += 6;
        mac32_per_s[form] = double(blocks) * threads * iters * 8.0 / (best * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    LSP_CUDA(ctx, cudaGetLastError());
    return LSP_OK;
}

// The fastest of the forms above: the denominator of the integer roofline.
extern "C" int lsp_int_peak(lsp_ctx* ctx, double* mac32_per_s) {
    if (!mac32_per_s) return LSP_ERR_PARAM;
    double v[LSP_INT_PEAK_FORMS];
    LSP_TRY(lsp_int_peaks(ctx, v));
    *mac32_per_s = v[0];
    for (int i = 1; i < LSP_INT_PEAK_FORMS; i++)
        if (v[i] > *mac32_per_s) *mac32_per_s = v[i];
    return LSP_OK;
}

extern "C" int lsp_set_poseidon2(lsp_ctx* ctx, int width, int sbox_d, int rounds_f, int rounds_p,
                                 const uint64_t* constants, const uint64_t* internal_diag_m1) {
    if (!ctx || !constants || !internal_diag_m1) return LSP_ERR_PARAM;
    if (width != 3) return set_err(ctx, LSP_ERR_PARAM, "only width 3 (Poseidon2Bls12337<3>) is supported, got %d", width);
    if (!(sbox_d == 3 || sbox_d == 5 || sbox_d == 7 || sbox_d == 11 || sbox_d == 17))
        return set_err(ctx, LSP_ERR_PARAM, "unsupported S-box degree %d", sbox_d);
    if (rounds_f <= 0 || (rounds_f & 1) || rounds_f / 2 > P2_MAX_HALF_F || rounds_p < 0 || rounds_p > P2_MAX_P)
        return set_err(ctx, LSP_ERR_PARAM, "unsupported round counts %d/%d", rounds_f, rounds_p);
    int n_const = rounds_f * 3 + rounds_p;
    for (int i = 0; i < n_const; i++)
        if (!limbs_reduced(constants + 4 * i)) return set_err(ctx, LSP_ERR_PARAM, "round constant %d is not reduced", i);
    for (int i = 0; i < 3; i++)
        if (!limbs_reduced(internal_diag_m1 + 4 * i)) return set_err(ctx, LSP_ERR_PARAM, "diag %d is not reduced", i);
    P2Params& p = ctx->p2;
    memset(&p, 0, sizeof p);
    p.half_f = rounds_f / 2;
    p.rounds_p = rounds_p;
    p.sbox_d = sbox_d;
    const uint64_t* c = constants;
    for (int r = 0; r < p.half_f; r++)
        for (int i = 0; i < 3; i++, c += 4) memcpy(&p.ext_initial[r][i], c, 32);
    for (int r = 0; r < p.half_f; r++)
        for (int i = 0; i < 3; i++, c += 4) memcpy(&p.ext_terminal[r][i], c, 32);
    for (int r = 0; r < rounds_p; r++, c += 4) memcpy(&p.internal[r], c, 32);
    memcpy(p.diag_m1, internal_diag_m1, 96);
    // diag (1,1,2) in Montgomery form -> add-only internal layer
    static const uint64_t ONE[4] = {0x7d1c7ffffffffff3ull, 0x7257f50f6ffffff2ull, 0x16d81575512c0feeull, 0x0d4bda322bbb9a9dull};
    static const uint64_t TWO[4] = {0xf0277fffffffffe5ull, 0x8b0573200fffffe3ull, 0xccfbddcc46206fdbull, 0x07ec4f05bd4a8fe3ull};
    p.diag_kind = (!memcmp(internal_diag_m1, ONE, 32) && !memcmp(internal_diag_m1 + 4, ONE, 32) &&
                   !memcmp(internal_diag_m1 + 8, TWO, 32))
                      ? 1
                      : 0;
    ctx->p2_set = true;
    return LSP_OK;
}

// ---------------------------------------------------------------------------
// parity probes
// ---------------------------------------------------------------------------
extern "C" int lsp_fr_op(lsp_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    if (!ctx || !a || !out || op < 0 || op > 4 || (op <= 2 && !b)) return LSP_ERR_PARAM;
    if (n == 0) return LSP_OK;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    Fr *da = nullptr, *db = nullptr, *dout = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&da, n * 32));
    LSP_TRY(tmp.get((void**)&db, n * 32));
    LSP_TRY(tmp.get((void**)&dout, n * 32));
    LSP_CUDA(ctx, cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (op <= 2) LSP_CUDA(ctx, cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_LAUNCH(ctx, k_fr_op, grid_for(ctx, n, 128), 128, 0, op, da, db, dout, n);
    LSP_CUDA(ctx, cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSP_OK;
}

extern "C" int lsp_poseidon2_permute(lsp_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n) {
    if (!ctx || !in || !out) return LSP_ERR_PARAM;
    if (!ctx->p2_set) return set_err(ctx, LSP_ERR_STATE, "lsp_set_poseidon2 has not been called");
    if (n == 0) return LSP_OK;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    Fr *din = nullptr, *dout = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&din, n * 96));
    LSP_TRY(tmp.get((void**)&dout, n * 96));
    LSP_CUDA(ctx, cudaMemcpyAsync(din, in, n * 96, cudaMemcpyHostToDevice, ctx->stream));
    LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_p2_permute<D>, grid_for(ctx, n, 128), 128, 0, ctx->p2, din, dout, n));
    LSP_CUDA(ctx, cudaMemcpyAsync(out, dout, n * 96, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSP_OK;
}

// ---------------------------------------------------------------------------
// matrices
// ---------------------------------------------------------------------------
int lsp::mat_alloc(lsp_ctx* ctx, size_t rows, size_t width, lsp_mat** out) {
    lsp_mat* m = new lsp_mat();
    m->rows = rows;
    m->width = width;
    int rc = dev_alloc(ctx, (void**)&m->d, rows * width * 32);
    if (rc != LSP_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return LSP_OK;
}

namespace lsp {
int rowmajor_to_colmajor(lsp_ctx* ctx, const Fr* rm_dev, size_t rows, size_t width, Fr* cm_dev) {
    LSP_LAUNCH(ctx, k_rm_to_cm, grid_for(ctx, rows * width, 256), 256, 0, rm_dev, cm_dev, rows, width, size_t(0), rows);
    return LSP_OK;
}
}  // namespace lsp

extern "C" int lsp_mat_upload(lsp_ctx* ctx, const uint64_t* rowmajor, size_t rows, size_t width, lsp_mat** out) {
    if (!ctx || !rowmajor || !out || rows == 0 || width == 0) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    lsp_mat* m = nullptr;
    LSP_TRY(mat_alloc(ctx, rows, width, &m));
    Fr* stage = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&stage, rows * width * 32));
    LSP_CUDA(ctx, cudaMemcpyAsync(stage, rowmajor, rows * width * 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_LAUNCH(ctx, k_rm_to_cm, grid_for(ctx, rows * width, 256), 256, 0, stage, m->d, rows, width, size_t(0), rows);
    *out = m;
    return LSP_OK;
}

extern "C" int lsp_mat_download_rows(lsp_ctx* ctx, const lsp_mat* m, size_t row0, size_t nrows, uint64_t* out) {
    if (!ctx || !m || !out || row0 + nrows > m->rows) return LSP_ERR_PARAM;
    if (nrows == 0) return LSP_OK;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    Fr* stage = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&stage, nrows * m->width * 32));
    LSP_LAUNCH(ctx, k_cm_to_rm, grid_for(ctx, nrows * m->width, 256), 256, 0, m->d, stage, m->rows, m->width, row0, nrows);
    LSP_CUDA(ctx, cudaMemcpyAsync(out, stage, nrows * m->width * 32, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSP_OK;
}

extern "C" int lsp_mat_download(lsp_ctx* ctx, const lsp_mat* m, uint64_t* out) {
    if (!m) return LSP_ERR_PARAM;
    return lsp_mat_download_rows(ctx, m, 0, m->rows, out);
}

extern "C" size_t lsp_mat_rows(const lsp_mat* m) { return m ? m->rows : 0; }
extern "C" size_t lsp_mat_width(const lsp_mat* m) { return m ? m->width : 0; }

extern "C" void lsp_mat_free(lsp_ctx* ctx, lsp_mat* m) {
    if (!m) return;
    if (m->owns && ctx) dev_free(ctx, m->d);
    delete m;
}

// ---------------------------------------------------------------------------
// Merkle MMCS
// ---------------------------------------------------------------------------
namespace lsp {

// One layer of 2-to-1 compressions.  Layers that cannot give every SM sub-partition a warp of
// one-thread-per-node work are latency-bound: they take the three-lanes-per-node kernel.
static int compress_layer(lsp_ctx* ctx, const Fr* in, Fr* out, size_t n_out) {
    if (n_out <= size_t(ctx->sm_count) * 4 * 10) {
        LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_compress_layer_tri<D>, unsigned((n_out + 9) / 10), 32, 0, ctx->p2, in, out, n_out));
    } else {
        LSP_DISPATCH_SBOX(ctx->p2.sbox_d,
                          LSP_LAUNCH(ctx, k_compress_layer<D>, grid_for(ctx, n_out, 128, LSP_P2_MINB), 128, 0, ctx->p2, in, out, n_out));
    }
    return LSP_OK;
}

// Levels n_out, n_out/2, ..., 1 in one cooperative launch (n_out <= 10 warps x 4 sub-partitions x SMs).
static int compress_top(lsp_ctx* ctx, const Fr* in, Fr* out, size_t n_out) {
    int n_levels = ilog2(n_out) + 1;
    if (n_levels == 1) return compress_layer(ctx, in, out, n_out);
    if (!ctx->grid_barrier) LSP_CUDA(ctx, cudaMalloc((void**)&ctx->grid_barrier, 4));
    LSP_CUDA(ctx, cudaMemsetAsync(ctx->grid_barrier, 0, 4, ctx->stream));
    unsigned grid = unsigned((n_out + 9) / 10);
    unsigned* bar = ctx->grid_barrier;
    void* args[] = {(void*)&ctx->p2, (void*)&in, (void*)&out, (void*)&n_out, (void*)&n_levels, (void*)&bar};
    lsp::timing_begin(ctx, "k_compress_top_tri<D>");
    cudaError_t e = cudaErrorInvalidValue;
    switch (ctx->p2.sbox_d) {
        case 3: e = cudaLaunchCooperativeKernel((void*)k_compress_top_tri<3>, grid, 32, args, 0, ctx->stream); break;
        case 5: e = cudaLaunchCooperativeKernel((void*)k_compress_top_tri<5>, grid, 32, args, 0, ctx->stream); break;
        case 7: e = cudaLaunchCooperativeKernel((void*)k_compress_top_tri<7>, grid, 32, args, 0, ctx->stream); break;
        case 11: e = cudaLaunchCooperativeKernel((void*)k_compress_top_tri<11>, grid, 32, args, 0, ctx->stream); break;
        case 17: e = cudaLaunchCooperativeKernel((void*)k_compress_top_tri<17>, grid, 32, args, 0, ctx->stream); break;
    }
    lsp::timing_end(ctx);
    ctx->launches++;
    LSP_CUDA(ctx, e);
    return LSP_OK;
}
static bool is_top(const lsp_ctx* ctx, size_t n_out) { return n_out <= size_t(ctx->sm_count) * 4 * 10 && is_pow2(n_out); }

// digests must hold 2h-1 elements; cols is a device array of `width` column pointers.
int merkle_build(lsp_ctx* ctx, const Fr* const* d_cols, int width, size_t h, Fr* digests) {
    if (!ctx->p2_set) return set_err(ctx, LSP_ERR_STATE, "lsp_set_poseidon2 has not been called");
    int log_h = ilog2(h);
    LSP_DISPATCH_SBOX(ctx->p2.sbox_d,
                      LSP_LAUNCH(ctx, k_leaf_hash<D>, grid_for(ctx, h, 128, LSP_P2_MINB), 128, 0, ctx->p2, d_cols, width, h, digests));
    for (int k = 0; k < log_h; k++) {
        size_t n_out = h >> (k + 1);
        const Fr* in = digests + tree_layer_offset(h, k);
        Fr* out = digests + tree_layer_offset(h, k + 1);
        if (is_top(ctx, n_out)) return compress_top(ctx, in, out, n_out);   // this level and every one above it
        LSP_TRY(compress_layer(ctx, in, out, n_out));
    }
    return LSP_OK;
}

// FRI commit-phase tree over a vector viewed as (len/2) rows of 2: the leaf digest of
// row j is hash_iter([v[2j], v[2j+1]]) = one permutation = the compression function,
// so the whole tree is compress layers applied to the vector itself.
int merkle_build_pairs(lsp_ctx* ctx, const Fr* vec, size_t len, Fr* digests) {
    if (!ctx->p2_set) return set_err(ctx, LSP_ERR_STATE, "lsp_set_poseidon2 has not been called");
    size_t h = len / 2;
    int log_h = ilog2(h);
    const Fr* in = vec;
    for (int k = 0; k <= log_h; k++) {
        size_t n_out = h >> k;
        Fr* out = digests + tree_layer_offset(h, k);
        if (is_top(ctx, n_out)) return compress_top(ctx, in, out, n_out);
        LSP_TRY(compress_layer(ctx, in, out, n_out));
        in = out;
    }
    return LSP_OK;
}

}  // namespace lsp

extern "C" int lsp_merkle_commit(lsp_ctx* ctx, const lsp_mat* const* mats, int n_mats, uint64_t root_out[4], lsp_tree** out) {
    if (!ctx || !mats || n_mats <= 0 || !out) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t h = mats[0]->rows;
    if (!is_pow2(h)) return set_err(ctx, LSP_ERR_PARAM, "matrix height %zu is not a power of two", h);
    std::vector<const Fr*> cols;
    for (int i = 0; i < n_mats; i++) {
        if (!mats[i] || mats[i]->rows != h)
            return set_err(ctx, LSP_ERR_PARAM, "mixed-height commits are not on the reference's path");
        for (size_t c = 0; c < mats[i]->width; c++) cols.push_back(mats[i]->d + c * h);
    }
    lsp_tree* t = new lsp_tree();
    t->height = h;
    t->log_h = ilog2(h);
    t->total_width = cols.size();
    for (int i = 0; i < n_mats; i++) t->mats.push_back(mats[i]);
    int rc = dev_alloc(ctx, (void**)&t->digests, (2 * h - 1) * 32);
    if (rc == LSP_OK) rc = dev_alloc(ctx, (void**)&t->d_cols, cols.size() * sizeof(Fr*));
    if (rc != LSP_OK) {
        lsp_tree_free(ctx, t);
        return rc;
    }
    // column-pointer table: small pageable copy, synchronous w.r.t. the host buffer
    LSP_CUDA(ctx, cudaMemcpyAsync(t->d_cols, cols.data(), cols.size() * sizeof(Fr*), cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rc = merkle_build(ctx, t->d_cols, int(cols.size()), h, t->digests);
    if (rc != LSP_OK) {
        lsp_tree_free(ctx, t);
        return rc;
    }
    if (root_out) {
        LSP_CUDA(ctx, cudaMemcpyAsync(root_out, t->digests + (2 * h - 2), 32, cudaMemcpyDeviceToHost, ctx->stream));
        LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *out = t;
    return LSP_OK;
}

extern "C" int lsp_merkle_open_batch(lsp_ctx* ctx, const lsp_tree* t, size_t index, uint64_t* rows_out, uint64_t* siblings_out) {
    if (!ctx || !t || !rows_out || !siblings_out || index >= t->height) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    Fr* buf = nullptr;
    size_t n = t->total_width + t->log_h;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&buf, n * 32));
    LSP_LAUNCH(ctx, k_gather_row, unsigned((t->total_width + 63) / 64), 64, 0, t->d_cols, int(t->total_width), index, buf);
    if (t->log_h > 0)
        LSP_LAUNCH(ctx, k_gather_siblings, 1, 64, 0, t->digests, t->height, t->log_h, index, buf + t->total_width);
    LSP_CUDA(ctx, cudaMemcpyAsync(rows_out, buf, t->total_width * 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (t->log_h > 0)
        LSP_CUDA(ctx, cudaMemcpyAsync(siblings_out, buf + t->total_width, size_t(t->log_h) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSP_OK;
}

extern "C" int lsp_merkle_layer(lsp_ctx* ctx, const lsp_tree* t, int layer, uint64_t* out) {
    if (!ctx || !t || !out || layer < 0 || layer > t->log_h) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = t->height >> layer;
    LSP_CUDA(ctx, cudaMemcpyAsync(out, t->digests + tree_layer_offset(t->height, layer), n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSP_OK;
}

extern "C" size_t lsp_merkle_height(const lsp_tree* t) { return t ? t->height : 0; }

extern "C" void lsp_tree_free(lsp_ctx* ctx, lsp_tree* t) {
    if (!t) return;
    if (ctx) {
        dev_free(ctx, t->digests);
        dev_free(ctx, (void*)t->d_cols);
    }
    delete t;
}

extern "C" int lsp_hash_rows(lsp_ctx* ctx, const uint64_t* rowmajor, size_t rows, size_t width, uint64_t* digests_out) {
    if (!ctx || !rowmajor || !digests_out || rows == 0) return LSP_ERR_PARAM;
    if (!ctx->p2_set) return set_err(ctx, LSP_ERR_STATE, "lsp_set_poseidon2 has not been called");
    if (width == 0) {  // hash_iter of an empty row is state[0] = 0 with no permutation
        memset(digests_out, 0, rows * 32);
        return LSP_OK;
    }
    lsp_mat* m = nullptr;
    LSP_TRY(lsp_mat_upload(ctx, rowmajor, rows, width, &m));
    std::vector<const Fr*> cols;
    for (size_t c = 0; c < width; c++) cols.push_back(m->d + c * rows);
    const Fr** d_cols = nullptr;
    Fr* dig = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&d_cols, cols.size() * sizeof(Fr*)));
    LSP_TRY(tmp.get((void**)&dig, rows * 32));
    LSP_CUDA(ctx, cudaMemcpyAsync(d_cols, cols.data(), cols.size() * sizeof(Fr*), cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_leaf_hash<D>, grid_for(ctx, rows, 128, LSP_P2_MINB), 128, 0, ctx->p2,
                                                 (const Fr* const*)d_cols, int(width), rows, dig));
    LSP_CUDA(ctx, cudaMemcpyAsync(digests_out, dig, rows * 32, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    lsp_mat_free(ctx, m);
    return LSP_OK;
}
