// STARK-specific kernels: device-resident Fiat-Shamir challenger, inverse
// denominators over the LDE coset, the fused permutation-AIR quotient kernel,
// opened values, and the FRI fold.
//
// Reference anchors (algorithms: published Plonky3 of the pinned era):
//   Challenger = HashChallenger<Val,Hash,1>      bin/src/config.rs:23, bin/src/main.rs:78
//   LineaAIR::eval -> eval_permutation           air/src/lib.rs:47-54,116-167
//   TwoAdicFriPcs::open / FRI commit phase       bin/src/config.rs:24-25, bin/src/main.rs:58-66
#include "stark.cuh"

using namespace lsp;

#ifndef LSP_GRIND_MINB
#define LSP_GRIND_MINB 8   // resident 128-thread blocks per SM, as for the other one-thread-per-permutation kernels
#endif

namespace lsp {

// ===========================================================================
// HashChallenger<Val,Hash,1> on the device.  The transcript is a serial chain of
// permutations by construction; keeping it on the device removes every host round
// trip from prove().  The hashing kernels run as ONE WARP (<<<1,32>>>) on the
// three-lanes-per-permutation shape (p2_permute_tri): a lone permutation costs 30
// S-box latencies instead of 46.  Every lane triple computes the same digest.
// ===========================================================================
template <int D>
__device__ __forceinline__ Fr ch_hash_input(const P2Params& P, const DevChallenger* ch) {
    // PaddingFreeSponge<Perm,3,2,1>::hash_iter(input_buffer); all 32 lanes call
    const int lane = threadIdx.x & 31, k = lane / 3, w = lane - 3 * k;
    Fr s = fr_zero();
    const int n = ch->n_input;
#pragma unroll 1
    for (int i = 0; i < n; i += 2) {
        if (w == 0) s = ch->input[i];
        if (w == 1 && i + 1 < n) s = ch->input[i + 1];  // odd tail: state[1] keeps its stale value
        p2_permute_tri<D>(P, s, w, 3 * k);
    }
    Fr out;
#pragma unroll
    for (int i = 0; i < 8; i++) out.l[i] = __shfl_sync(0xffffffffu, s.l[i], 0);
    return out;
}

// sample(): output_buffer is always empty when sample is called on this path
// (OUT_LEN = 1 and every flush is immediately popped), so sample == flush + pop.
template <int D>
__device__ __forceinline__ Fr ch_sample(const P2Params& P, DevChallenger* ch) {
    Fr out = ch_hash_input<D>(P, ch);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        ch->input[0] = out;
        ch->n_input = 1;
    }
    __syncwarp();
    return out;
}

__device__ __forceinline__ void ch_observe(DevChallenger* ch, const Fr& v) {
    if (ch->n_input < CH_CAP)
        ch->input[ch->n_input++] = v;
    else
        ch->overflow = 1;
}

__global__ void k_ch_init(DevChallenger* ch) {
    ch->n_input = 0;
    ch->overflow = 0;
}
__global__ void k_ch_observe(DevChallenger* ch, const Fr* vals, int n) {
    for (int i = 0; i < n; i++) ch_observe(ch, fr_load(vals + i));
}
template <int D>
__global__ void __launch_bounds__(32) k_ch_sample(const __grid_constant__ P2Params P, DevChallenger* ch, Fr* out) {
    Fr v = ch_sample<D>(P, ch);
    if (threadIdx.x == 0) fr_store(out, v);
}
// observe(vals[0..n)) then sample() in ONE launch, optionally copying the observed values to `copy_to` first (a
// commitment on its way into the proof): what every commit-phase round does with its root.  Saves two launches and a
// device copy per round on the latency chain of the FRI commit phase.
template <int D>
__global__ void __launch_bounds__(32) k_ch_observe_sample(const __grid_constant__ P2Params P, DevChallenger* ch, const __grid_constant__ ObserveList L,
                                                          Fr* copy_to, Fr* out) {
    if (threadIdx.x == 0)
        for (int s = 0; s < L.n_seg; s++)
            for (int i = 0; i < L.n[s]; i++) {
                const Fr v = fr_load(L.p[s] + i);
                ch_observe(ch, v);
                if (copy_to && s == 0) fr_store(copy_to + i, v);
            }
    __syncwarp();
    Fr v = ch_sample<D>(P, ch);
    if (threadIdx.x == 0) fr_store(out, v);
}
// canonical integer of a Montgomery-form element: multiply by 1
__device__ __forceinline__ Fr fr_from_mont(const Fr& a) {
    Fr one = fr_zero();
    one.l[0] = 1;
    return fr_mul(a, one);
}
template <int D>
__global__ void __launch_bounds__(32) k_ch_sample_bits(const __grid_constant__ P2Params P, DevChallenger* ch, int bits, int n,
                                                       uint32_t* idx) {
#pragma unroll 1
    for (int q = 0; q < n; q++) {
        Fr c = fr_from_mont(ch_sample<D>(P, ch));
        if (threadIdx.x == 0) idx[q] = bits >= 32 ? c.l[0] : (c.l[0] & ((1u << bits) - 1u));
    }
}

// grind: every trial hashes [input_buffer..., witness] (a clone of the challenger observing the witness, then
// sample_bits).  All trials share every sponge block but the last, so one warp absorbs those once (k_ch_grind_prefix) and
// a trial is ONE permutation; each thread tests the witness base + its global index, and the smallest passing witness
// of the first chunk that has one wins, which makes the result deterministic (the reference's rayon find_any is not).
struct GrindPrefix {
    Fr s[3];     // sponge state after the blocks that do not hold the witness
    Fr last;     // input[n-1]: shares the witness' block when n is odd
    int n_odd;
};
template <int D>
__global__ void __launch_bounds__(32) k_ch_grind_prefix(const __grid_constant__ P2Params P, const DevChallenger* ch, GrindPrefix* out) {
    const int lane = threadIdx.x & 31, k = lane / 3, w = lane - 3 * k;
    const int n = ch->n_input, full = n & ~1;   // the witness is element n: blocks (0,1) .. (full-2, full-1) come first
    Fr s = fr_zero();
#pragma unroll 1
    for (int i = 0; i < full; i += 2) {
        if (w < 2) s = ch->input[i + w];
        p2_permute_tri<D>(P, s, w, 3 * k);
    }
    if (lane < 3) out->s[lane] = s;
    if (lane == 0) {
        out->n_odd = n & 1;
        out->last = (n & 1) ? ch->input[n - 1] : fr_zero();
    }
}
template <int D>
__global__ void __launch_bounds__(128, LSP_GRIND_MINB) k_ch_grind_chunk(const __grid_constant__ P2Params P, const GrindPrefix* __restrict__ pre, int bits,
                                                                        unsigned long long base, unsigned long long* best) {
    unsigned long long w = base + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    // witness as a field element: canonical w -> Montgomery
    Fr wf = fr_zero();
    wf.l[0] = uint32_t(w);
    wf.l[1] = uint32_t(w >> 32);
    wf = fr_mul(wf, fr_const(FR_R2));
    // the witness' block: [input[n-1], witness] when the buffer held an odd count, else [witness] alone with the
    // second word stale (overwrite-mode sponge)
    LSP_P2_SLOT_DECL(128);
    Fr s0 = pre->n_odd ? pre->last : wf, s1 = pre->n_odd ? wf : pre->s[1], s2 = pre->s[2];
    p2_permute<D, 128>(P, s0, s1, s2, LSP_P2_SLOT(128));
    Fr c = fr_from_mont(s0);
    uint32_t low = bits >= 32 ? c.l[0] : (c.l[0] & ((1u << bits) - 1u));
    if (low == 0) atomicMin(best, w);
}
template <int D>
__global__ void __launch_bounds__(32) k_ch_apply_witness(const __grid_constant__ P2Params P, DevChallenger* ch,
                                                         const unsigned long long* best, Fr* witness_out) {
    unsigned long long w = *best;
    Fr wf = fr_zero();
    wf.l[0] = uint32_t(w);
    wf.l[1] = uint32_t(w >> 32);
    wf = fr_mul(wf, fr_const(FR_R2));
    if (threadIdx.x == 0) {
        fr_store(witness_out, wf);
        ch_observe(ch, wf);    // check_witness: observe(witness) ...
    }
    __syncwarp();
    (void)ch_sample<D>(P, ch);  // ... then sample_bits(bits) consumes one sample
}

// The verifier's whole transcript in one launch (one warp): every challenge `verify` + `TwoAdicFriPcs::verify`
// draw, in their order.  Nothing here depends on the query data, so the
// per-query kernels can all start from its outputs.
template <int D>
__global__ void __launch_bounds__(32) k_verify_transcript(const __grid_constant__ P2Params P, DevChallenger* ch,
                                                          const __grid_constant__ VerifyTranscriptArgs A) {
    const bool lead = threadIdx.x == 0;
    auto observe = [&](const Fr* v, int n) {
        if (lead)
            for (int i = 0; i < n; i++) ch_observe(ch, fr_load(v + i));
        __syncwarp();
    };
    auto sample_to = [&](Fr* out) {
        Fr v = ch_sample<D>(P, ch);
        if (lead) fr_store(out, v);
    };
    auto sample_bits = [&](int bits) {
        Fr c = fr_from_mont(ch_sample<D>(P, ch));
        return bits >= 32 ? c.l[0] : (c.l[0] & ((1u << bits) - 1u));
    };
    if (lead) {
        ch->n_input = 0;
        ch->overflow = 0;
        Fr ln = fr_zero();
        ln.l[0] = uint32_t(A.log_n);
        ch_observe(ch, fr_mul(ln, fr_const(FR_R2)));      // observe(degree_bits)
    }
    __syncwarp();
    observe(A.trace_commit, 1);
    observe(A.publics, 2);
    sample_to(A.scal + VT_ALPHA);
    observe(A.quot_commit, 1);
    sample_to(A.scal + VT_ZETA);
    if (A.alpha_before_openings) sample_to(A.scal + VT_ALPHA_FRI);   // fork-era order: batching challenge first, values never observed
    if (A.observe_opened_values) observe(A.opened, A.n_opened);
    if (!A.alpha_before_openings) sample_to(A.scal + VT_ALPHA_FRI);
    for (int r = 0; r < A.n_rounds; r++) {
        observe(A.fri_commits + r, 1);
        sample_to(A.betas + r);
    }
    observe(A.final_poly, A.n_final);
    observe(A.pow_witness, 1);                              // check_witness: observe, then sample_bits == 0
    uint32_t pow_low = sample_bits(A.pow_bits);
    if (lead) *A.pow_low = pow_low;
    for (int q = 0; q < A.n_queries; q++) {
        uint32_t i = sample_bits(A.log_l);
        if (lead) A.idx[q] = i;
    }
}

int challenger_init(lsp_ctx* ctx, DevChallenger* ch) {
    LSP_LAUNCH(ctx, k_ch_init, 1, 1, 0, ch);
    return LSP_OK;
}
int challenger_observe_dev(lsp_ctx* ctx, DevChallenger* ch, const Fr* vals, int n) {
    LSP_LAUNCH(ctx, k_ch_observe, 1, 1, 0, ch, vals, n);
    return LSP_OK;
}
int challenger_sample(lsp_ctx* ctx, DevChallenger* ch, Fr* out_dev) {
    LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_ch_sample<D>, 1, 32, 0, ctx->p2, ch, out_dev));
    return LSP_OK;
}
int challenger_observe_sample(lsp_ctx* ctx, DevChallenger* ch, const ObserveList& L, Fr* copy_to, Fr* out_dev) {
    LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_ch_observe_sample<D>, 1, 32, 0, ctx->p2, ch, L, copy_to, out_dev));
    return LSP_OK;
}
int challenger_observe_sample(lsp_ctx* ctx, DevChallenger* ch, const Fr* vals, int n, Fr* copy_to, Fr* out_dev) {
    ObserveList L;
    L.n_seg = 1;
    L.p[0] = vals;
    L.n[0] = n;
    return challenger_observe_sample(ctx, ch, L, copy_to, out_dev);
}
int challenger_sample_bits(lsp_ctx* ctx, DevChallenger* ch, int bits, int n, uint32_t* idx_out) {
    LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_ch_sample_bits<D>, 1, 32, 0, ctx->p2, ch, bits, n, idx_out));
    return LSP_OK;
}
int verify_transcript(lsp_ctx* ctx, DevChallenger* ch, const VerifyTranscriptArgs& A) {
    LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_verify_transcript<D>, 1, 32, 0, ctx->p2, ch, A));
    return LSP_OK;
}
int challenger_grind(lsp_ctx* ctx, DevChallenger* ch, int bits, Fr* witness_out) {
    unsigned long long* best = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&best, 8));
    if (bits == 0) {
        LSP_CUDA(ctx, cudaMemsetAsync(best, 0, 8, ctx->stream));  // witness 0 always passes
    } else {
        // expected 2^bits trials; test chunks until one contains a witness (the host reads one u64 per chunk)
        GrindPrefix* pre = nullptr;
        LSP_TRY(tmp.get((void**)&pre, sizeof(GrindPrefix)));
        LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_ch_grind_prefix<D>, 1, 32, 0, ctx->p2, (const DevChallenger*)ch, pre));
        // ~2 expected witnesses per chunk (a chunk without one: e^-2), between one wave of the device and 2^25
        const unsigned long long chunk = 1ull << (bits + 1 < 17 ? 17 : (bits + 1 > 25 ? 25 : bits + 1));
        unsigned long long h_best = ~0ull;
        for (unsigned long long base = 0;; base += chunk) {
            LSP_CUDA(ctx, cudaMemsetAsync(best, 0xff, 8, ctx->stream));
            LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_ch_grind_chunk<D>, unsigned(chunk / 128), 128, 0, ctx->p2,
                                                         (const GrindPrefix*)pre, bits, base, best));
            LSP_CUDA(ctx, cudaMemcpyAsync(&h_best, best, 8, cudaMemcpyDeviceToHost, ctx->stream));
            LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (h_best != ~0ull) break;
            if (base > (1ull << 44)) return set_err(ctx, LSP_ERR_STATE, "grind: no witness found for %d bits", bits);
        }
    }
    LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_ch_apply_witness<D>, 1, 32, 0, ctx->p2, ch, (const unsigned long long*)best, witness_out));
    return LSP_OK;
}

// ===========================================================================
// AIR config upload
// ===========================================================================
int upload_air_cfgs(lsp_ctx* ctx, const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* cfgs, int n_cfgs,
                    size_t width, PermCfgDev* out, void** blob) {
    if (n_lookups < 0 || n_cfgs < 0 || n_lookups + n_cfgs <= 0 || (n_lookups && !lookups) || (n_cfgs && !cfgs))
        return set_err(ctx, LSP_ERR_PARAM, "no AIR configs");
    std::vector<uint32_t> h;
    size_t w_sum = 0;
    // ---- lookups: [lk_off x n_lookups][records...]
    std::vector<uint32_t> lk_off(n_lookups), lk;
    auto in_range = [&](uint32_t id) { return size_t(id) < width; };
    for (int i = 0; i < n_lookups; i++) {
        const lsp_lookup_air_cfg& c = lookups[i];
        if (c.n_a_cols == 0 || c.n_tables == 0 || c.n_b_cols == 0 || !c.a_ids || !c.b_ids || !c.b_filter_ids || !c.b_inverses_ids ||
            !c.occurrences_ids)
            return set_err(ctx, LSP_ERR_PARAM, "lookup config %d is incomplete", i);
        bool ok = in_range(c.a_filter_id) && in_range(c.a_inverses_id) && in_range(c.check_id);
        for (uint32_t k = 0; k < c.n_a_cols; k++) ok = ok && in_range(c.a_ids[k]);
        for (uint32_t t = 0; t < c.n_tables; t++) {
            ok = ok && in_range(c.b_filter_ids[t]) && in_range(c.b_inverses_ids[t]) && in_range(c.occurrences_ids[t]);
            for (uint32_t k = 0; k < c.n_b_cols; k++) ok = ok && in_range(c.b_ids[size_t(t) * c.n_b_cols + k]);
        }
        if (!ok) return set_err(ctx, LSP_ERR_PARAM, "lookup config %d: column id out of range", i);
        lk_off[i] = uint32_t(lk.size());
        lk.insert(lk.end(), {c.n_a_cols, c.n_tables, c.n_b_cols, c.a_filter_id, c.a_inverses_id, c.check_id});
        lk.insert(lk.end(), c.a_ids, c.a_ids + c.n_a_cols);
        for (uint32_t t = 0; t < c.n_tables; t++) {
            lk.insert(lk.end(), {c.b_filter_ids[t], c.b_inverses_ids[t], c.occurrences_ids[t]});
            lk.insert(lk.end(), c.b_ids + size_t(t) * c.n_b_cols, c.b_ids + size_t(t + 1) * c.n_b_cols);
        }
        w_sum += size_t(c.n_a_cols) + size_t(c.n_tables) * (size_t(c.n_b_cols) + 3) + 3;  // air/src/air_lookup.rs:37-39
    }
    // ---- permutations: [n_cols x n][ids_off x n][b_inv x n][check x n][ids...]
    size_t total_ids = 0;
    for (int i = 0; i < n_cfgs; i++) {
        const lsp_perm_air_cfg& c = cfgs[i];
        if (c.n_cols == 0 || !c.a_ids || !c.b_ids) return set_err(ctx, LSP_ERR_PARAM, "AIR config %d has no columns", i);
        for (uint32_t k = 0; k < c.n_cols; k++)
            if (c.a_ids[k] >= width || c.b_ids[k] >= width) return set_err(ctx, LSP_ERR_PARAM, "AIR config %d: column id out of range", i);
        if (c.b_inverse_id >= width || c.check_id >= width) return set_err(ctx, LSP_ERR_PARAM, "AIR config %d: column id out of range", i);
        total_ids += 2 * size_t(c.n_cols);
        w_sum += 2 * size_t(c.n_cols) + 2;  // air/src/air_permutation.rs:21-23
    }
    if (w_sum != width) return set_err(ctx, LSP_ERR_PARAM, "AIR width %zu != trace width %zu", w_sum, width);
    const size_t perm_words = 4 * size_t(n_cfgs) + total_ids;
    h.resize(perm_words + size_t(n_lookups) + lk.size());
    size_t off = 0;
    for (int i = 0; i < n_cfgs; i++) {
        h[i] = cfgs[i].n_cols;
        h[n_cfgs + i] = uint32_t(off);
        h[2 * n_cfgs + i] = cfgs[i].b_inverse_id;
        h[3 * n_cfgs + i] = cfgs[i].check_id;
        for (uint32_t k = 0; k < cfgs[i].n_cols; k++) h[4 * n_cfgs + off + k] = cfgs[i].a_ids[k];
        for (uint32_t k = 0; k < cfgs[i].n_cols; k++) h[4 * n_cfgs + off + cfgs[i].n_cols + k] = cfgs[i].b_ids[k];
        off += 2 * size_t(cfgs[i].n_cols);
    }
    for (int i = 0; i < n_lookups; i++) h[perm_words + i] = lk_off[i];
    for (size_t i = 0; i < lk.size(); i++) h[perm_words + n_lookups + i] = lk[i];
    uint32_t* d = nullptr;
    LSP_TRY(dev_alloc(ctx, (void**)&d, h.size() * 4));
    LSP_CUDA(ctx, cudaMemcpyAsync(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->n_cfgs = n_cfgs;
    out->n_cols = d;
    out->ids_off = d + n_cfgs;
    out->b_inverse_id = d + 2 * n_cfgs;
    out->check_id = d + 3 * n_cfgs;
    out->ids = d + 4 * n_cfgs;
    out->n_lookups = n_lookups;
    out->lk_off = d + perm_words;
    out->lk = d + perm_words + n_lookups;
    *blob = d;
    return LSP_OK;
}

int upload_perm_cfgs(lsp_ctx* ctx, const lsp_perm_air_cfg* cfgs, int n_cfgs, size_t width, PermCfgDev* out, void** blob) {
    if (!cfgs || n_cfgs <= 0) return set_err(ctx, LSP_ERR_PARAM, "no AIR configs");
    return upload_air_cfgs(ctx, nullptr, 0, cfgs, n_cfgs, width, out, blob);
}

// ===========================================================================
// Inverse denominators  E[p] = 1 / (g * w^{bitrev(p)} - z),  w = omega_{2^m}.
//
// 1/(x-z) over the whole coset costs ~2.6 multiplications per point and ONE
// field inversion, by descending the squaring tree of the coset: with
// u = z/g, u_l = u^(2^(m-l)), v a 2^(l+1)-th root of unity,
//     1/(v - u_{l+1}) = (v + u_{l+1}) / (v^2 - u_l),   1/(-v - u_{l+1}) = (u_{l+1} - v) / (v^2 - u_l)
// so level l+1 is one multiplication per node away from level l, and the root is
// the scalar 1/(1 - u^(2^m)).  In bit-reversed order the children of node j
// are 2j and 2j+1, which is exactly the storage order of the committed LDE.
// ===========================================================================
constexpr int INVDEN_MAX_LEVELS = 40;
struct InvDenScalars {
    Fr u[INVDEN_MAX_LEVELS];  // u[l] = (z/g)^(2^(m-l)), l = 0..m
    Fr e0;                    // g^-1 / (1 - u[0])   (the g^-1 makes E the inverse of g*w - z directly)
};

__global__ void k_invden_setup(const FieldConsts* __restrict__ fc, const Fr* __restrict__ z, int m, InvDenScalars* __restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= gridDim.x * blockDim.x) return;
    Fr ginv = fr_load(&fc->gen_inv);
    Fr u = fr_mul(fr_load(z + p), ginv);
    InvDenScalars* o = out + p;
    o->u[m] = u;
    for (int l = m - 1; l >= 0; l--) {
        u = fr_sqr(u);
        o->u[l] = u;
    }
    Fr d = fr_sub(fr_one(), u);
    o->e0 = fr_mul(ginv, fr_inv(d));
}

// node (level l, index j) -> factor taking E_l[j>>1... ] to E_{l}[j]: uses v = w_{2^l}^{bitrev_{l-1}(j>>1)}
__device__ __forceinline__ Fr invden_step(const Fr& e_parent, const InvDenScalars* S, const Fr* __restrict__ tw, int m, int l_child,
                                          size_t j_child) {
    // parent level l = l_child-1, parent index jp = j_child>>1
    int l = l_child - 1;
    size_t jp = j_child >> 1;
    Fr v = fr_load_nc(tw + (size_t(bitrev32(uint32_t(jp), l)) << (m - l - 1)));
    Fr u = S->u[l_child];
    Fr f = (j_child & 1) ? fr_sub(u, v) : fr_add(v, u);
    return fr_mul(e_parent, f);
}

// One thread per node of level `la`; walks down from the root.
__global__ void __launch_bounds__(128) k_invden_top(const InvDenScalars* __restrict__ S, const Fr* __restrict__ tw, int m, int la,
                                                    Fr* __restrict__ out) {
    size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (j >= (size_t(1) << la)) return;
    Fr e = S->e0;
    for (int l = 1; l <= la; l++) e = invden_step(e, S, tw, m, l, j >> (la - l));
    fr_store(out + j, e);
}

// From level `la` values down to the leaves (level m): each thread owns one node of
// level lb = m - 3 and writes its 8 leaves.
__global__ void __launch_bounds__(128) k_invden_expand(const InvDenScalars* __restrict__ S, const Fr* __restrict__ tw, int m, int la,
                                                       int lb, const Fr* __restrict__ top, Fr* __restrict__ out, size_t j0, size_t n_nodes) {
    size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (j >= n_nodes) return;
    j += j0;
    out -= j0 << (m - lb);  // out[0] is leaf j0 << depth
    Fr e[8];
    e[0] = fr_load_nc(top + (j >> (lb - la)));
    for (int l = la + 1; l <= lb; l++) e[0] = invden_step(e[0], S, tw, m, l, j >> (lb - l));
    const int depth = m - lb;  // <= 3
    for (int d = 0; d < depth; d++) {
        for (int i = (1 << d) - 1; i >= 0; i--) {
            size_t jc = ((j << d) + i) << 1;
            Fr par = e[i];
            e[2 * i + 1] = invden_step(par, S, tw, m, lb + d + 1, jc + 1);
            e[2 * i] = invden_step(par, S, tw, m, lb + d + 1, jc);
        }
    }
    for (int i = 0; i < (1 << depth); i++) fr_store(out + (j << depth) + i, e[i]);
}

int inverse_denominators_range(lsp_ctx* ctx, const Fr* z_dev, int n_points, int log_m, size_t p0, size_t count, Fr* const* out) {
    if (log_m < 1 || log_m >= INVDEN_MAX_LEVELS) return set_err(ctx, LSP_ERR_PARAM, "inverse_denominators: 2^%d unsupported", log_m);
    const Fr* tw = nullptr;
    LSP_TRY(twiddles(ctx, log_m, false, &tw));
    InvDenScalars* S = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&S, sizeof(InvDenScalars) * n_points));
    LSP_LAUNCH(ctx, k_invden_setup, 1, n_points, 0, (const FieldConsts*)ctx->fc, z_dev, log_m, S);
    int la = log_m < 10 ? log_m : 10;
    int lb = log_m - 3 > la ? log_m - 3 : la;
    const int depth = log_m - lb;
    const bool direct = la == log_m;
    if (!direct && ((p0 | count) & ((size_t(1) << depth) - 1)))
        return set_err(ctx, LSP_ERR_PARAM, "inverse_denominators: range must be aligned to %d rows", 1 << depth);
    Fr* top = nullptr;
    LSP_TRY(tmp.get((void**)&top, (size_t(1) << la) * 32));
    for (int p = 0; p < n_points; p++) {
        LSP_LAUNCH(ctx, k_invden_top, unsigned(((size_t(1) << la) + 127) / 128), 128, 0, S + p, tw, log_m, la, top);
        if (direct) {
            LSP_CUDA(ctx, cudaMemcpyAsync(out[p], top + p0, count * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
            size_t n_nodes = count >> depth;
            LSP_LAUNCH(ctx, k_invden_expand, unsigned((n_nodes + 127) / 128), 128, 0, S + p, tw, log_m, la, lb, (const Fr*)top, out[p],
                       p0 >> depth, n_nodes);
        }
    }
    return LSP_OK;
}

int inverse_denominators(lsp_ctx* ctx, const Fr* z_dev, int n_points, int log_m, Fr* const* out) {
    return inverse_denominators_range(ctx, z_dev, n_points, log_m, 0, size_t(1) << log_m, out);
}

// ===========================================================================
// Fused quotient kernel for the permutation AIR.
// One thread per storage row p of the first N*q rows of the committed trace LDE
// (= evaluations over g*H_{Nq} in bit-reversed order).  Natural index
// i = bitrev(p); `next` is natural index i+q, i.e. storage row bitrev(i+q).
// ===========================================================================
struct QuotientArgs {
    const Fr* lde;        // column-major, column stride = lde_rows
    const Fr* lde_next;   // optional: the neighbouring sub-coset's rows (same shape); nullptr = next rows are in `lde`
    int next_log_m;       // with lde_next: log2 of the sub-coset's size, or -1 when the next row is the same local row
    size_t lde_rows;
    int log_n, log_q;
    PermCfgDev cfg;
    const Fr* publics;    // [alpha_air, delta]
    const Fr* alpha;      // STARK folding challenge
    const Fr* inv_first;  // 1/(x - 1)          per storage row
    const Fr* inv_last;   // 1/(x - w_N^-1)     per storage row
    const Fr* zh;         // per chunk c: Z_H = g^N * w_q^c - 1, then 1/Z_H   (2q entries)
    const Fr* tw_nq;      // omega_{Nq}^j, j < Nq/2
    const Fr* w_n_inv;    // omega_N^-1
    const FieldConsts* fc;
    Fr* chunks;           // q columns of N
    size_t p0, count;     // storage rows [p0, p0+count) handled by this launch; lde/inv_* are indexed by p - p_base
    size_t p_base;        // storage row that lde[0] / inv_first[0] correspond to
};

__global__ void k_quotient_setup(const FieldConsts* __restrict__ fc, int log_n, int log_q, Fr* __restrict__ zh, Fr* __restrict__ w_n_inv, Fr* __restrict__ pts) {
    int c = threadIdx.x;
    int q = 1 << log_q;
    if (c < q) {
        Fr g = fr_load(&fc->gen);
        Fr gn = g;
        for (int i = 0; i < log_n; i++) gn = fr_sqr(gn);
        Fr wq = fr_pow_u32(fr_two_adic_generator(fc, log_q), uint32_t(c));
        Fr z = fr_sub(fr_mul(gn, wq), fr_one());
        fr_store(zh + c, z);
        fr_store(zh + q + c, fr_inv(z));
    }
    if (c == 0) {
        Fr w = fr_two_adic_generator(fc, log_n);
        Fr wi = fr_pow_u32(w, uint32_t((size_t(1) << log_n) - 1));  // w^(N-1) = w^-1
        fr_store(w_n_inv, wi);
        fr_store(pts, fr_one());      // selector points: 1 and w_N^-1
        fr_store(pts + 1, wi);
    }
}

__global__ void __launch_bounds__(128) k_quotient_permutation(const __grid_constant__ QuotientArgs A) {
    const int lnq = A.log_n + A.log_q;
    const size_t nq = size_t(1) << lnq;
    const uint32_t q = 1u << A.log_q;
    const Fr alpha_air = fr_load(A.publics), delta = fr_load(A.publics + 1), alpha = fr_load(A.alpha);
    const Fr w_n_inv = fr_load(A.w_n_inv);
    const Fr gen = fr_load(&A.fc->gen);
    const Fr one = fr_one();
    for (size_t pi = blockIdx.x * size_t(blockDim.x) + threadIdx.x; pi < A.count; pi += size_t(gridDim.x) * blockDim.x) {
        const size_t pg = A.p0 + pi;                 // global storage row
        const uint32_t i = bitrev32(uint32_t(pg), lnq);
        const uint32_t i_next = (i + q) & uint32_t(nq - 1);
        const long long p = (long long)(pg - A.p_base);   // local row
        long long pn;
        if (!A.lde_next) {
            pn = (long long)(size_t(bitrev32(i_next, lnq)) - A.p_base);   // same N-row block
        } else {                                                          // second matrix: same point, or the following one
            long long r = p;
            if (A.next_log_m >= 0) r = bitrev32((bitrev32(uint32_t(p), A.next_log_m) + 1u) & ((1u << A.next_log_m) - 1u), A.next_log_m);
            pn = r + (long long)(A.lde_next - A.lde);
        }
        const uint32_t c = i & (q - 1);
        // x = g * w_{Nq}^i
        Fr wi = (i < nq / 2) ? fr_load_nc(A.tw_nq + i) : fr_neg(fr_load_nc(A.tw_nq + (i - nq / 2)));
        if (nq == 1) wi = one;
        Fr x = fr_mul(gen, wi);
        Fr zh = fr_load_nc(A.zh + c), zh_inv = fr_load_nc(A.zh + q + c);
        Fr is_first = fr_mul(zh, fr_load_nc(A.inv_first + p));
        Fr is_last = fr_mul(zh, fr_load_nc(A.inv_last + p));
        Fr is_trans = fr_sub(x, w_n_inv);
        Fr acc = fold_air_constraints(A.cfg, A.lde, A.lde_rows, p, pn, alpha_air, delta, alpha, is_first, is_last, is_trans);
        Fr qv = fr_mul(acc, zh_inv);
        // split_evals: chunk c holds natural rows c, c+q, ...
        fr_store(A.chunks + (size_t(c) << A.log_n) + (i >> A.log_q), qv);
    }
}

// Quotient values for storage rows [p0, p0+count) of the quotient domain (count a multiple of N:
// whole chunks).  `lde` points at storage row p_base of every column (column stride lde_rows);
// chunk c = bitrev(block) is written to chunks + c*N.
int quotient_permutation_range(lsp_ctx* ctx, const Fr* lde, size_t lde_rows, size_t p_base, int log_n, int log_q, const PermCfgDev& cfg,
                               const Fr* publics_dev, const Fr* alpha_dev, size_t p0, size_t count, Fr* chunks, const Fr* lde_next,
                               bool next_rotated) {
    int lnq = log_n + log_q;
    size_t nq = size_t(1) << lnq;
    if (lnq < 1) return set_err(ctx, LSP_ERR_PARAM, "trace of height 1 with a single quotient chunk is unsupported");
    const size_t whole = (size_t(1) << log_n) - 1;
    if (p0 + count > nq || (!lde_next && ((count | p0) & whole)))
        return set_err(ctx, LSP_ERR_PARAM, "quotient range must cover whole chunks unless the next rows are supplied");
    // The selector tables -- Z_H per chunk, 1/(x - 1) and 1/(x - w_N^-1) over the quotient domain -- depend on the
    // shape only, not on the trace or the challenges: built once per (log_n, log_q, row range) and kept with the
    // context (2 * count * 32 bytes), like the twiddles.  Saves two inverse-denominator sweeps and a Fermat
    // inversion per prove (quotient stage 0.96 -> ~0.45 ms at 2^19 rows).
    size_t q = size_t(1) << log_q;
    const auto key = std::make_tuple(log_n, log_q, p0, count);
    auto it = ctx->quot_sel.find(key);
    if (it == ctx->quot_sel.end()) {
        lsp_ctx::QuotSel sel;
        LSP_CUDA(ctx, cudaMalloc((void**)&sel.scal, (2 * q + 3) * 32));
        LSP_CUDA(ctx, cudaMalloc((void**)&sel.inv0, count * 32));
        LSP_CUDA(ctx, cudaMalloc((void**)&sel.inv1, count * 32));
        Fr* zh0 = sel.scal;
        LSP_LAUNCH(ctx, k_quotient_setup, 1, unsigned(q < 32 ? 32 : q), 0, (const FieldConsts*)ctx->fc, log_n, log_q, zh0, zh0 + 2 * q, zh0 + 2 * q + 1);
        Fr* inv01[2] = {sel.inv0, sel.inv1};
        LSP_TRY(inverse_denominators_range(ctx, zh0 + 2 * q + 1, 2, lnq, p0, count, inv01));
        it = ctx->quot_sel.emplace(key, sel).first;
    }
    Fr* scal = it->second.scal;  // zh[2q], w_n_inv, pts[2]
    Fr *zh = scal, *w_n_inv = scal + 2 * q;
    Fr* inv[2] = {it->second.inv0, it->second.inv1};
    const Fr* tw = nullptr;
    LSP_TRY(twiddles(ctx, lnq, false, &tw));
    QuotientArgs A;
    A.lde = lde;
    A.lde_next = lde_next;
    A.next_log_m = (lde_next && next_rotated) ? ilog2(count) : -1;
    A.lde_rows = lde_rows;
    A.log_n = log_n;
    A.log_q = log_q;
    A.cfg = cfg;
    A.publics = publics_dev;
    A.alpha = alpha_dev;
    A.inv_first = inv[0] - (p0 - p_base);  // indexed by local row p = pg - p_base
    A.inv_last = inv[1] - (p0 - p_base);
    A.zh = zh;
    A.tw_nq = tw;
    A.w_n_inv = w_n_inv;
    A.fc = ctx->fc;
    A.chunks = chunks;
    A.p0 = p0;
    A.count = count;
    A.p_base = p_base;
    LSP_LAUNCH(ctx, k_quotient_permutation, grid_for(ctx, count, 128), 128, 0, A);
    return LSP_OK;
}

int quotient_permutation(lsp_ctx* ctx, const Fr* lde, size_t lde_rows, int log_n, int log_q, const PermCfgDev& cfg,
                         const Fr* publics_dev, const Fr* alpha_dev, Fr* chunks) {
    size_t nq = size_t(1) << (log_n + log_q);
    if (nq > lde_rows) return set_err(ctx, LSP_ERR_PARAM, "quotient domain (2^%d) exceeds the LDE (%zu rows)", log_n + log_q, lde_rows);
    return quotient_permutation_range(ctx, lde, lde_rows, 0, log_n, log_q, cfg, publics_dev, alpha_dev, 0, nq, chunks);
}

// ===========================================================================
// Opened values from coefficient form:  y[c] = sum_k coeffs[c][k] z^k.
// (Same field element as the reference's barycentric `interpolate_coset`.)
// ===========================================================================
constexpr int EVAL_LO_BITS = 10;
__global__ void __launch_bounds__(128) k_point_pow_tables(const Fr* __restrict__ z, int log_n, int lo_bits, Fr* __restrict__ lo,
                                                          Fr* __restrict__ hi) {
    size_t n_lo = size_t(1) << lo_bits, n_hi = size_t(1) << (log_n - lo_bits);
    Fr base = fr_load(z);
    for (size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x; j < n_lo + n_hi; j += size_t(gridDim.x) * blockDim.x) {
        if (j < n_lo)
            fr_store(lo + j, fr_pow_u32(base, uint32_t(j)));
        else
            fr_store(hi + (j - n_lo), fr_pow_u32(base, uint32_t((j - n_lo) << lo_bits)));
    }
}

constexpr int EVAL_COLS = 4;      // columns per thread (register budget)
constexpr int EVAL_THREADS = 128;
// grid: (blocks over k, column groups).  partial[(group*gridDim.x + block)*EVAL_COLS + c]
__global__ void __launch_bounds__(EVAL_THREADS) k_eval_partial(const Fr* __restrict__ coeffs, size_t n, int width, int lo_bits,
                                                               const Fr* __restrict__ lo, const Fr* __restrict__ hi,
                                                               Fr* __restrict__ partial) {
    __shared__ uint4 red[EVAL_THREADS * 2];
    const int c0 = blockIdx.y * EVAL_COLS;
    Fr acc[EVAL_COLS];
#pragma unroll
    for (int c = 0; c < EVAL_COLS; c++) acc[c] = fr_zero();
    const size_t lo_mask = (size_t(1) << lo_bits) - 1;
    for (size_t k = blockIdx.x * size_t(blockDim.x) + threadIdx.x; k < n; k += size_t(gridDim.x) * blockDim.x) {
        Fr zp = fr_mul(fr_load_nc(lo + (k & lo_mask)), fr_load_nc(hi + (k >> lo_bits)));
#pragma unroll
        for (int c = 0; c < EVAL_COLS; c++)
            if (c0 + c < width) acc[c] = fr_add(acc[c], fr_mul(fr_load_nc(coeffs + size_t(c0 + c) * n + k), zp));
    }
    // block tree reduction, one column at a time
    for (int c = 0; c < EVAL_COLS; c++) {
        red[threadIdx.x] = make_uint4(acc[c].l[0], acc[c].l[1], acc[c].l[2], acc[c].l[3]);
        red[EVAL_THREADS + threadIdx.x] = make_uint4(acc[c].l[4], acc[c].l[5], acc[c].l[6], acc[c].l[7]);
        __syncthreads();
        for (int s = EVAL_THREADS / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                uint4 a0 = red[threadIdx.x], a1 = red[EVAL_THREADS + threadIdx.x];
                uint4 b0 = red[threadIdx.x + s], b1 = red[EVAL_THREADS + threadIdx.x + s];
                Fr a, b;
                a.l[0] = a0.x; a.l[1] = a0.y; a.l[2] = a0.z; a.l[3] = a0.w; a.l[4] = a1.x; a.l[5] = a1.y; a.l[6] = a1.z; a.l[7] = a1.w;
                b.l[0] = b0.x; b.l[1] = b0.y; b.l[2] = b0.z; b.l[3] = b0.w; b.l[4] = b1.x; b.l[5] = b1.y; b.l[6] = b1.z; b.l[7] = b1.w;
                a = fr_add(a, b);
                red[threadIdx.x] = make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]);
                red[EVAL_THREADS + threadIdx.x] = make_uint4(a.l[4], a.l[5], a.l[6], a.l[7]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            uint4 a0 = red[0], a1 = red[EVAL_THREADS];
            Fr a;
            a.l[0] = a0.x; a.l[1] = a0.y; a.l[2] = a0.z; a.l[3] = a0.w; a.l[4] = a1.x; a.l[5] = a1.y; a.l[6] = a1.z; a.l[7] = a1.w;
            fr_store(partial + (size_t(blockIdx.y) * gridDim.x + blockIdx.x) * EVAL_COLS + c, a);
        }
        __syncthreads();
    }
}
// One block per column: the n_blocks partial sums are added by 128 threads and a shared-memory tree (a single thread
// walking them one dependent global load at a time took 0.28 ms per call, four calls per prove).
__global__ void __launch_bounds__(EVAL_THREADS) k_eval_finish(const Fr* __restrict__ partial, int n_blocks, int width, Fr* __restrict__ y) {
    __shared__ uint4 red[EVAL_THREADS * 2];
    const int c = blockIdx.x;
    if (c >= width) return;
    const int grp = c / EVAL_COLS, cc = c % EVAL_COLS;
    Fr acc = fr_zero();
    for (int b = threadIdx.x; b < n_blocks; b += EVAL_THREADS) acc = fr_add(acc, fr_load_nc(partial + (size_t(grp) * n_blocks + b) * EVAL_COLS + cc));
    red[threadIdx.x] = make_uint4(acc.l[0], acc.l[1], acc.l[2], acc.l[3]);
    red[EVAL_THREADS + threadIdx.x] = make_uint4(acc.l[4], acc.l[5], acc.l[6], acc.l[7]);
    __syncthreads();
    for (int s = EVAL_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            uint4 a0 = red[threadIdx.x], a1 = red[EVAL_THREADS + threadIdx.x];
            uint4 b0 = red[threadIdx.x + s], b1 = red[EVAL_THREADS + threadIdx.x + s];
            Fr a, b;
            a.l[0] = a0.x; a.l[1] = a0.y; a.l[2] = a0.z; a.l[3] = a0.w; a.l[4] = a1.x; a.l[5] = a1.y; a.l[6] = a1.z; a.l[7] = a1.w;
            b.l[0] = b0.x; b.l[1] = b0.y; b.l[2] = b0.z; b.l[3] = b0.w; b.l[4] = b1.x; b.l[5] = b1.y; b.l[6] = b1.z; b.l[7] = b1.w;
            a = fr_add(a, b);
            red[threadIdx.x] = make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]);
            red[EVAL_THREADS + threadIdx.x] = make_uint4(a.l[4], a.l[5], a.l[6], a.l[7]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint4 a0 = red[0], a1 = red[EVAL_THREADS];
        Fr a;
        a.l[0] = a0.x; a.l[1] = a0.y; a.l[2] = a0.z; a.l[3] = a0.w; a.l[4] = a1.x; a.l[5] = a1.y; a.l[6] = a1.z; a.l[7] = a1.w;
        fr_store(y + c, a);
    }
}

int eval_columns_at(lsp_ctx* ctx, const Fr* coeffs, size_t n, size_t width, const Fr* z_dev, Fr* y_dev) {
    int log_n = ilog2(n);
    int lo_bits = log_n < EVAL_LO_BITS ? log_n : EVAL_LO_BITS;
    size_t n_lo = size_t(1) << lo_bits, n_hi = size_t(1) << (log_n - lo_bits);
    Fr* tab = nullptr;
    Scratch tmp(ctx);
    LSP_TRY(tmp.get((void**)&tab, (n_lo + n_hi) * 32));
    LSP_LAUNCH(ctx, k_point_pow_tables, unsigned((n_lo + n_hi + 127) / 128), 128, 0, z_dev, log_n, lo_bits, tab, tab + n_lo);
    int groups = int((width + EVAL_COLS - 1) / EVAL_COLS);
    int blocks = int((n + EVAL_THREADS - 1) / EVAL_THREADS);
    int cap = ctx->sm_count * 4;
    if (blocks > cap) blocks = cap;
    Fr* partial = nullptr;
    LSP_TRY(tmp.get((void**)&partial, size_t(groups) * blocks * EVAL_COLS * 32));
    LSP_LAUNCH(ctx, k_eval_partial, dim3(blocks, groups), EVAL_THREADS, 0, coeffs, n, int(width), lo_bits, (const Fr*)tab,
               (const Fr*)(tab + n_lo), partial);
    LSP_LAUNCH(ctx, k_eval_finish, unsigned(width), EVAL_THREADS, 0, (const Fr*)partial, blocks, int(width), y_dev);
    return LSP_OK;
}

// ===========================================================================
// FRI fold (arity 2):  out[j] = (1/2 + t_j) in[2j] + (1/2 - t_j) in[2j+1],
//                      t_j = (beta/2) * w_len^{-bitrev(j)}
// written as  (lo+hi)/2 + t_j (lo-hi)  -> 2 multiplications per pair.
// ===========================================================================
__global__ void __launch_bounds__(128) k_fri_fold(const Fr* __restrict__ in, size_t h, int log_h, const Fr* __restrict__ beta,
                                                  const Fr* __restrict__ tw_inv /* w_len^-e, e < len/2 */, Fr* __restrict__ out, size_t j0) {
    // h = pairs handled here; j0 = global index of the first one; log_h = log2 of the GLOBAL pair count
    Fr half_beta = fr_halve(fr_load(beta));
    for (size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x; j < h; j += size_t(gridDim.x) * blockDim.x) {
        Fr lo = fr_load_nc(in + 2 * j), hi = fr_load_nc(in + 2 * j + 1);
        Fr t = fr_mul(half_beta, fr_load_nc(tw_inv + bitrev32(uint32_t(j0 + j), log_h)));
        Fr r = fr_add(fr_halve(fr_add(lo, hi)), fr_mul(t, fr_sub(lo, hi)));
        fr_store(out + j, r);
    }
}

// Folds the slice [2*j0, 2*j0 + 2*h_local) of a vector of global length `len`; `in`/`out` point at the slice.
int fri_fold_range(lsp_ctx* ctx, const Fr* in, size_t len, size_t j0, size_t h_local, const Fr* beta_dev, Fr* out) {
    int log_len = ilog2(len);
    const Fr* tw = nullptr;
    LSP_TRY(twiddles(ctx, log_len, true, &tw));
    LSP_LAUNCH(ctx, k_fri_fold, grid_for(ctx, h_local, 128), 128, 0, in, h_local, log_len - 1, beta_dev, tw, out, j0);
    return LSP_OK;
}
int fri_fold(lsp_ctx* ctx, const Fr* in, size_t len, const Fr* beta_dev, Fr* out) {
    return fri_fold_range(ctx, in, len, 0, len / 2, beta_dev, out);
}

}  // namespace lsp
