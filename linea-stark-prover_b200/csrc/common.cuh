// Shared host-side declarations of the CUDA library (context, handles, launch helpers).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/lsp_b200.h"
#include "fr.cuh"
#include "poseidon2.cuh"

namespace lsp {
struct TwiddleCache;
// Field parameters the fork-only `p3-bls12-377-fr` crate fixes: Montgomery form, device resident.
struct FieldConsts {
    Fr gen;      // Val::GENERATOR: the coset shift of TwoAdicFriPcs
    Fr gen_inv;  // 1 / gen
    Fr root47;   // two_adic_generator(47)
};
}

struct lsp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    bool p2_set = false;
    lsp::P2Params p2;
    lsp::FieldConsts* fc = nullptr;      // device memory; lsp_set_field_consts
    // TwoAdicFriPcs::open of the pinned fork samples the batching challenge BEFORE computing the opened values and never
    // observes them; later upstream observes them first.  lsp_set_transcript_flags selects either.
    bool alpha_before_openings = true;
    bool observe_opened_values = false;
    uint64_t launches = 0;
    // omega_{2^k}^j tables, j < 2^(k-1), forward and inverse, keyed by k
    std::map<int, lsp::Fr*> tw_fwd, tw_inv;
    // pinned staging for small D2H reads
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // optional per-kernel timing (bench.py / profiles): CUDA events around every launch
    unsigned* grid_barrier = nullptr;   // arrival counter of the fused tree-top kernel
    bool ntt_smem_opt_in = false;       // k_ntt_tile's > 48 KiB of dynamic shared memory enabled on this device
    // shape-only selector tables of the quotient kernel, keyed by (log_n, log_q, first row, rows)
    struct QuotSel {
        lsp::Fr *scal = nullptr, *inv0 = nullptr, *inv1 = nullptr;
    };
    std::map<std::tuple<int, int, size_t, size_t>, QuotSel> quot_sel;
    bool timing = false;
    bool timing_leaf_only = false;   // mode 2: only the Poseidon2 leaf-hash launches (two events per commit)
    bool timing_armed = false;       // set by timing_begin when it recorded e0 for the current launch
    const char* phase = "";
    struct TimingRec {
        const char* name;
        const char* phase;
        cudaEvent_t e0, e1;
    };
    std::vector<TimingRec> timing_recs;
    std::vector<cudaEvent_t> ev_pool;
};

// Column-major ("planar") on the device: column c is the contiguous run
// d[c*rows .. (c+1)*rows).  Host matrices are row-major; upload/download transpose.
struct lsp_mat {
    lsp::Fr* d = nullptr;
    size_t rows = 0, width = 0;
    bool owns = true;
};

struct lsp_tree {
    lsp::Fr* digests = nullptr;  // layer k at offset 2h - (2h >> k), 2h-1 digests in total
    size_t height = 0;
    int log_h = 0;
    std::vector<const lsp_mat*> mats;
    size_t total_width = 0;
    const lsp::Fr** d_cols = nullptr;  // device array of total_width column base pointers
};

namespace lsp {

inline int set_err(lsp_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        ctx->err = buf;
    }
    return code;
}

#define LSP_CUDA(ctx, call)                                                                     \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return lsp::set_err(ctx, e__ == cudaErrorMemoryAllocation ? LSP_ERR_NOMEM : LSP_ERR_CUDA, \
                                "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

#define LSP_TRY(call)            \
    do {                         \
        int rc__ = (call);       \
        if (rc__ != LSP_OK) return rc__; \
    } while (0)

inline cudaEvent_t timing_event(lsp_ctx* ctx) {
    cudaEvent_t e;
    if (!ctx->ev_pool.empty()) {
        e = ctx->ev_pool.back();
        ctx->ev_pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
inline void timing_begin(lsp_ctx* ctx, const char* name) {
    ctx->timing_armed = false;
    if (!ctx->timing) return;
    if (ctx->timing_leaf_only && !strstr(name, "k_leaf_hash")) return;
    lsp_ctx::TimingRec r{name, ctx->phase, timing_event(ctx), timing_event(ctx)};
    cudaEventRecord(r.e0, ctx->stream);
    ctx->timing_recs.push_back(r);
    ctx->timing_armed = true;
}
inline void timing_end(lsp_ctx* ctx) {
    if (!ctx->timing_armed) return;
    cudaEventRecord(ctx->timing_recs.back().e1, ctx->stream);
    ctx->timing_armed = false;
}

// Launch on the ctx stream, count it, and surface launch-time errors.
#define LSP_LAUNCH(ctx, kernel, grid, block, smem, ...)                      \
    do {                                                                     \
        lsp::timing_begin(ctx, #kernel);                                     \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);     \
        lsp::timing_end(ctx);                                                \
        (ctx)->launches++;                                                   \
        LSP_CUDA(ctx, cudaGetLastError());                                   \
    } while (0)

inline size_t tree_layer_offset(size_t h, int k) { return 2 * h - ((2 * h) >> k); }

inline int dev_alloc(lsp_ctx* ctx, void** p, size_t bytes) {
    LSP_CUDA(ctx, cudaMallocAsync(p, bytes ? bytes : 16, ctx->stream));
    return LSP_OK;
}
inline void dev_free(lsp_ctx* ctx, void* p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

// Temporary device buffers of one call: everything handed out is returned to the stream-ordered pool when the
// holder goes out of scope, on the success path and on every early return alike.
struct Scratch {
    lsp_ctx* ctx;
    std::vector<void*> ptrs;
    explicit Scratch(lsp_ctx* c) : ctx(c) {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    ~Scratch() {
        for (void* p : ptrs) dev_free(ctx, p);
    }
    int get(void** p, size_t bytes) {
        int rc = dev_alloc(ctx, p, bytes);
        if (rc == LSP_OK) ptrs.push_back(*p);
        return rc;
    }
};

// A matrix under construction: freed on every early return, handed over with release().
struct MatGuard {
    lsp_ctx* ctx;
    lsp_mat* m;
    ~MatGuard() {
        if (m) lsp_mat_free(ctx, m);
    }
    lsp_mat* release() {
        lsp_mat* r = m;
        m = nullptr;
        return r;
    }
};

inline int ilog2(size_t n) {
    int k = 0;
    while ((size_t(1) << k) < n) k++;
    return k;
}
inline bool is_pow2(size_t n) { return n && !(n & (n - 1)); }

// Grid size for a grid-stride kernel: enough CTAs to fill every SM several times over.
inline unsigned grid_for(const lsp_ctx* ctx, size_t n, unsigned block, unsigned ctas_per_sm = 8) {
    size_t need = (n + block - 1) / block;
    size_t cap = size_t(ctx->sm_count) * ctas_per_sm;
    if (need < 1) need = 1;
    return unsigned(need < cap ? need : cap);
}

// ---- internal kernels' host entry points (defined in the .cu files) -------
int mat_alloc(lsp_ctx* ctx, size_t rows, size_t width, lsp_mat** out);
int twiddles(lsp_ctx* ctx, int log_n, bool inverse, const Fr** out);

}  // namespace lsp
