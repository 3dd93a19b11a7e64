// Device verifier: `p3_uni_stark::verify` over `TwoAdicFriPcs::verify`, the call the reference's
// `main` makes right after `prove` (bin/src/main.rs:88-96), plus `Mmcs::verify_batch`
// (bin/src/config.rs:19-20) on its own.
//
// Verification is a few thousand permutations, so the shape is latency, not throughput:
//   1. k_verify_transcript (stark.cu)  one warp replays the whole Fiat-Shamir transcript: alpha, zeta,
//                                      the FRI batching challenge, every beta, the proof-of-work check
//                                      and all query indices.  Nothing in it depends on query data.
//   2. k_verify_fold_ood               one thread per query: index check, reduced opening at the queried
//                                      point (two inversions), the fold chain through every round (the
//                                      round points are a squaring chain, no inversion), final-polynomial
//                                      check.  Writes the leaf pair of every commit-phase opening.
//   3. k_verify_path_jobs + k_merkle_paths_tri   three lanes per (query, tree) Merkle path -- trace, quotient,
//                                      one per FRI round: 33 x 21 independent paths at the baseline shape.
//      (last block of launch 2)        quotient recombination and the AIR constraints at zeta, through the same
//                                      fold_air_constraints the quotient kernel uses; every inversion on its own thread.
// The host then reads one small status block and reports the FIRST failing check in the
// reference verifier's order, so a rejected proof yields the same reason as the CPU verifier.
//
// Proof layout: see prover.cu.  The stored per-query index is redundant (the verifier samples
// it); a proof whose stored index differs is rejected as malformed.
#include "../csrc/stark.cuh"

using namespace lsp;

namespace lsp {

struct PathJob {
    const Fr* row;    // leaf row, `width` elements
    const Fr* sibs;   // log_h siblings, leaf level first
    const Fr* root;
    int* bad;         // set to 1 when the recomputed root differs
    uint32_t index;
    int width, log_h;
};

// MerkleTreeMmcs::verify_batch for one matrix: hash the row (PaddingFreeSponge, rate 2), then compress with the
// siblings up to the root -- on three lanes per path (p2_permute_tri: 30 S-box latencies per permutation instead of
// 46).  Sponge steps and tree levels share ONE inlined permutation site.  A warp carries ten paths; every lane runs as
// many permutations as the longest of them, shorter paths keep their result aside.
template <int D>
__global__ void __launch_bounds__(32) k_merkle_paths_tri(const __grid_constant__ P2Params P, const PathJob* __restrict__ jobs, int n_jobs) {
    const int lane = threadIdx.x, k = lane / 3, w = lane - 3 * k;
    const int t = blockIdx.x * 10 + k;
    const bool live = k < 10 && t < n_jobs;
    const PathJob J = jobs[live ? t : 0];
    const int n_leaf = (J.width + 1) / 2, total = live ? n_leaf + J.log_h : 0;
    const int longest = __reduce_max_sync(0xffffffffu, total);
    Fr s = fr_zero(), result = fr_zero();
#pragma unroll 1
    for (int step = 0; step < longest; step++) {
        // word 0 of the triple's state, for the compress steps (all lanes take part in the shuffle)
        Fr node;
#pragma unroll
        for (int i = 0; i < 8; i++) node.l[i] = __shfl_sync(0xffffffffu, s.l[i], 3 * k < 30 ? 3 * k : 30);
        if (step < n_leaf) {
            const int c = 2 * step + w;                         // w = 2: the capacity word carries over
            if (w < 2 && c < J.width) s = fr_load(J.row + c);   // odd tail: word 1 keeps its stale value
        } else if (step < total) {
            const int lvl = step - n_leaf;
            const Fr sib = fr_load(J.sibs + lvl);
            const bool right = (J.index >> lvl) & 1u;
            s = w == 2 ? fr_zero() : ((w == 0) == right ? sib : node);
        }
        p2_permute_tri<D>(P, s, w, 3 * k < 30 ? 3 * k : 30);
        if (step == total - 1) result = s;
    }
    if (live && w == 0) {
        const Fr root = fr_load(J.root);
        bool same = true;
#pragma unroll
        for (int i = 0; i < 8; i++) same = same && (root.l[i] == result.l[i]);
        *J.bad = same ? 0 : 1;
    }
}

// Deserialisation check: every field element of the proof must be the canonical representative (< r), as
// `Bls12_377Fr`'s deserialiser demands; the per-query index words are small integers and pass trivially.
__global__ void __launch_bounds__(128) k_verify_canonical(const Fr* __restrict__ proof, size_t n_elems, int* __restrict__ bad) {
    const uint32_t p[8] = {LSP_P0, LSP_P1, LSP_P2, LSP_P3, LSP_P4, LSP_P5, LSP_P6, LSP_P7};
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n_elems; i += size_t(gridDim.x) * blockDim.x) {
        Fr v = fr_load(proof + i);
        uint32_t t[8];
        if (u256_sub(t, v.l, p) == 0) atomicExch(bad, 1);   // no borrow: v >= r
    }
}

// A proof rebuilt from its serialised form (lsp_proof_deserialize) does not carry the per-query indices -- `Proof` has no
// such field, the verifier samples them: those slots hold the all-ones marker and receive the sampled index here.
__global__ void k_adopt_sampled_indices(Fr* __restrict__ queries, size_t per_query, const uint32_t* __restrict__ idx, int n_queries) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= n_queries) return;
    Fr* slot = queries + size_t(qi) * per_query;
    const Fr v = fr_load(slot);
    bool marker = true;
#pragma unroll
    for (int i = 0; i < 8; i++) marker = marker && v.l[i] == 0xffffffffu;
    if (marker) {
        Fr r = fr_zero();
        r.l[0] = idx[qi];
        fr_store(slot, r);
    }
}

struct VerifyArgs {
    const Fr *proof, *p_local, *p_next, *p_chunks, *p_commits, *p_final, *p_queries;
    const Fr *publics, *scal, *betas;
    const uint32_t* idx;
    int W, q, log_n, log_q, log_l, n_rounds, n_final, n_queries;
    size_t per_query;
    Fr* ev;        // n_queries x n_rounds x 2: leaf rows of the commit-phase openings
    int* status;   // n_queries x (4 + n_rounds): [index, final poly, trace path, quotient path, round paths...]
    int* ood_bad;
    PermCfgDev cfg;
    const FieldConsts* fc;
};
constexpr int ST_INDEX = 0, ST_FINAL = 1, ST_TRACE = 2, ST_QUOT = 3, ST_ROUND = 4;

__device__ __forceinline__ bool fr_same(const Fr& a, const Fr& b) {
    bool s = true;
#pragma unroll
    for (int i = 0; i < 8; i++) s = s && (a.l[i] == b.l[i]);
    return s;
}

// verify_query for one query: open_input's reduced opening, then the fold chain.
__device__ __forceinline__ void verify_fold_one(const VerifyArgs& A, int qi) {
    if (qi >= A.n_queries) return;
    const uint32_t index = A.idx[qi];
    const Fr* in = A.p_queries + size_t(qi) * A.per_query;
    int* st = A.status + size_t(qi) * (ST_ROUND + A.n_rounds);
    {   // the stored index is a raw integer and must be the sampled one
        const Fr stored = fr_load(in);
        bool ok = stored.l[0] == index;
#pragma unroll
        for (int i = 1; i < 8; i++) ok = ok && stored.l[i] == 0;
        st[ST_INDEX] = ok ? 0 : 1;
    }
    const Fr* trow = in + 1;
    const Fr* qrow = in + 1 + A.W + A.log_l;
    const Fr zeta = fr_load(A.scal + VT_ZETA), a = fr_load(A.scal + VT_ALPHA_FRI);
    const Fr zeta_next = fr_mul(zeta, fr_two_adic_generator(A.fc, A.log_n));
    // s = w_L^bitrev(index): the queried point is x = g * s
    const Fr wl = fr_two_adic_generator(A.fc, A.log_l);
    const uint32_t e = bitrev32(index, A.log_l);
    Fr s = fr_pow_u32(wl, e);
    Fr s_inv = fr_pow_u32(wl, uint32_t(((size_t(1) << A.log_l) - e) & ((size_t(1) << A.log_l) - 1)));
    const Fr x = fr_mul(fr_load(&A.fc->gen), s);
    const Fr ix0 = fr_inv(fr_sub(x, zeta)), ix1 = fr_inv(fr_sub(x, zeta_next));
    Fr ap = fr_one(), ro = fr_zero();
    for (int c = 0; c < A.W; c++) {
        ro = fr_add(ro, fr_mul(ap, fr_mul(fr_sub(fr_load(trow + c), fr_load(A.p_local + c)), ix0)));
        ap = fr_mul(ap, a);
    }
    for (int c = 0; c < A.W; c++) {
        ro = fr_add(ro, fr_mul(ap, fr_mul(fr_sub(fr_load(trow + c), fr_load(A.p_next + c)), ix1)));
        ap = fr_mul(ap, a);
    }
    for (int c = 0; c < A.q; c++) {
        ro = fr_add(ro, fr_mul(ap, fr_mul(fr_sub(fr_load(qrow + c), fr_load(A.p_chunks + c)), ix0)));
        ap = fr_mul(ap, a);
    }
    // Fold chain.  In round r the element `di` of the round's domain sits at the point s_r (s_0 = s,
    // s_{r+1} = s_r^2); the pair's even member is x0 = +-s_r, and 1/(x1 - x0) = -(1/2) / x0.
    const Fr neg_half = fr_neg(fr_const(FR_HALF));
    Fr folded = fr_zero();
    uint32_t di = index;
    size_t o = size_t(1) + A.W + A.log_l + A.q + A.log_l;
    Fr* ev = A.ev + size_t(qi) * A.n_rounds * 2;
    for (int r = 0; r < A.n_rounds; r++) {
        if (r == 0) folded = fr_add(folded, ro);   // the only reduced opening on this path enters at the top height
        const Fr sib = fr_load(in + o);
        const bool odd = di & 1u;
        const Fr ev0 = odd ? sib : folded, ev1 = odd ? folded : sib;
        fr_store(ev + 2 * r, ev0);
        fr_store(ev + 2 * r + 1, ev1);
        const Fr x0 = odd ? fr_neg(s) : s, x0_inv = odd ? fr_neg(s_inv) : s_inv;
        const Fr beta = fr_load(A.betas + r);
        folded = fr_add(ev0, fr_mul(fr_mul(fr_sub(beta, x0), fr_sub(ev1, ev0)), fr_mul(neg_half, x0_inv)));
        di >>= 1;
        s = fr_sqr(s);
        s_inv = fr_sqr(s_inv);
        o += size_t(A.log_l - r);  // sibling value + (log_l - 1 - r) digests
    }
    // final polynomial at the remaining point (s after n_rounds squarings)
    Fr acc = fr_zero(), xp = fr_one();
    for (int j = 0; j < A.n_final; j++) {
        acc = fr_add(acc, fr_mul(fr_load(A.p_final + j), xp));
        xp = fr_mul(xp, s);
    }
    st[ST_FINAL] = fr_same(acc, folded) ? 0 : 1;
}

// One job per (query, tree): trace, quotient, then one per commit-phase round.
__global__ void k_verify_path_jobs(const __grid_constant__ VerifyArgs A, PathJob* __restrict__ jobs) {
    const int per = 2 + A.n_rounds;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.n_queries * per) return;
    const int qi = t / per, path = t - qi * per;
    const uint32_t index = A.idx[qi];
    const Fr* in = A.p_queries + size_t(qi) * A.per_query;
    int* st = A.status + size_t(qi) * (ST_ROUND + A.n_rounds);
    PathJob J;
    if (path == 0) {
        J.row = in + 1;
        J.width = A.W;
        J.sibs = in + 1 + A.W;
        J.log_h = A.log_l;
        J.index = index;
        J.root = A.proof;
        J.bad = st + ST_TRACE;
    } else if (path == 1) {
        J.row = in + 1 + A.W + A.log_l;
        J.width = A.q;
        J.sibs = J.row + A.q;
        J.log_h = A.log_l;
        J.index = index;
        J.root = A.proof + 1;
        J.bad = st + ST_QUOT;
    } else {
        const int r = path - 2;
        size_t o = size_t(1) + A.W + A.log_l + A.q + A.log_l + size_t(r) * A.log_l - size_t(r) * (r - 1) / 2;
        J.row = A.ev + (size_t(qi) * A.n_rounds + r) * 2;
        J.width = 2;
        J.sibs = in + o + 1;
        J.log_h = A.log_l - 1 - r;
        J.index = index >> (r + 1);
        J.root = A.p_commits + r;
        J.bad = st + ST_ROUND + r;
    }
    jobs[t] = J;
}

// Out-of-domain check: sum_i zp_i(zeta) * chunk_i(zeta) == folded_constraints(zeta) / Z_H(zeta).
// One block of VERIFY_BLOCK threads.  Every inversion the identity needs runs on its own thread -- the q(q-1) cross factors
// of the zp_i on threads (i, j), 1/(zeta - 1), 1/(zeta - w_N^-1) and 1/Z_H(zeta) on three more -- so the depth is ONE
// Fermat inversion; 1/shift_j is a power of the inverse generators, not an inversion.  Thread 0 finishes.
constexpr int VERIFY_BLOCK = 96;
__device__ __forceinline__ void verify_ood_block(const VerifyArgs& A) {
    __shared__ Fr factor[64];
    __shared__ Fr inv3[3];   // 1/(zeta - 1), 1/(zeta - w_N^-1), 1/Z_H(zeta)
    const int t = threadIdx.x, q = A.q, i = t / q, j = t - i * q;
    const Fr zeta = fr_load(A.scal + VT_ZETA);
    const Fr one = fr_one();
    if (t < q * q && i != j) {
        const int lnq = A.log_n + A.log_q;
        const Fr wnq = fr_two_adic_generator(A.fc, lnq);
        const uint32_t mask = uint32_t((size_t(1) << lnq) - 1);
        // shift_k = g * w_{Nq}^k:  zeta / shift_j = zeta * g^-1 * w^-j,   shift_i / shift_j = w^(i-j)
        Fr a = fr_mul(fr_mul(zeta, fr_load(&A.fc->gen_inv)), fr_pow_u32(wnq, (0u - uint32_t(j)) & mask));
        Fr b = fr_pow_u32(wnq, (uint32_t(i) - uint32_t(j)) & mask);
        for (int k = 0; k < A.log_n; k++) {
            a = fr_sqr(a);
            b = fr_sqr(b);
        }
        factor[t] = fr_mul(fr_sub(a, one), fr_inv(fr_sub(b, one)));
    }
    if (t >= 64 && t < 67) {
        const Fr wn_inv = fr_pow_u32(fr_two_adic_generator(A.fc, A.log_n), uint32_t((size_t(1) << A.log_n) - 1));
        Fr v;
        if (t == 64) v = fr_sub(zeta, one);
        else if (t == 65) v = fr_sub(zeta, wn_inv);
        else {
            Fr zn = zeta;
            for (int k = 0; k < A.log_n; k++) zn = fr_sqr(zn);
            v = fr_sub(zn, one);
        }
        inv3[t - 64] = fr_inv(v);
    }
    __syncthreads();
    if (t != 0) return;
    Fr quotient = fr_zero();
    for (int ii = 0; ii < q; ii++) {
        Fr zp = one;
        for (int jj = 0; jj < q; jj++)
            if (jj != ii) zp = fr_mul(zp, factor[ii * q + jj]);
        quotient = fr_add(quotient, fr_mul(zp, fr_load(A.p_chunks + ii)));
    }
    Fr zn = zeta;
    for (int k = 0; k < A.log_n; k++) zn = fr_sqr(zn);
    const Fr z_h = fr_sub(zn, one);
    const Fr wn_inv = fr_pow_u32(fr_two_adic_generator(A.fc, A.log_n), uint32_t((size_t(1) << A.log_n) - 1));
    const Fr is_first = fr_mul(z_h, inv3[0]);
    const Fr is_last = fr_mul(z_h, inv3[1]);
    const Fr is_trans = fr_sub(zeta, wn_inv);
    const Fr folded = fold_air_constraints(A.cfg, A.p_local, 1, 0, size_t(A.W), fr_load(A.publics), fr_load(A.publics + 1),
                                           fr_load(A.scal + VT_ALPHA), is_first, is_last, is_trans);
    *A.ood_bad = fr_same(fr_mul(folded, inv3[2]), quotient) ? 0 : 1;
}

// Fold chains and the out-of-domain identity depend on the transcript only, not on each other: one launch, the last
// block takes the identity, the others one query per thread.
__global__ void __launch_bounds__(VERIFY_BLOCK) k_verify_fold_ood(const __grid_constant__ VerifyArgs A) {
    if (blockIdx.x + 1 == gridDim.x)
        verify_ood_block(A);
    else
        verify_fold_one(A, int(blockIdx.x) * VERIFY_BLOCK + int(threadIdx.x));
}

}  // namespace lsp

extern "C" int lsp_verify_air(lsp_ctx* ctx, const lsp_fri_config* fri, uint32_t log_n, size_t width, const lsp_lookup_air_cfg* lookups,
                              int n_lookups, const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                              const uint64_t* proof_words_in, size_t proof_words, float* device_ms_out) {
    if (!ctx || !fri || !publics || !proof_words_in || n_lookups < 0 || n_cfgs < 0 || n_lookups + n_cfgs <= 0) return LSP_ERR_PARAM;
    if ((n_cfgs && !cfgs) || (n_lookups && !lookups)) return LSP_ERR_PARAM;
    if (!ctx->p2_set) return set_err(ctx, LSP_ERR_STATE, "lsp_set_poseidon2 has not been called");
    const int log_q = lsp_air_log_quotient_degree_cfg(lookups, n_lookups, cfgs, n_cfgs), q = 1 << log_q;
    const int log_b = int(fri->log_blowup), log_l = int(log_n) + log_b;
    LSP_TRY(check_fri_config(ctx, fri, int(log_n), log_q));
    const int n_rounds = int(log_n) - int(fri->log_final_poly_len), log_f = log_b + int(fri->log_final_poly_len);
    // the proof's shape is a function of the parameters: any other length is InvalidProofShape
    if (proof_words != lsp_proof_words(log_n, uint32_t(width), uint32_t(log_q), fri)) return LSP_VERIFY_INVALID_PROOF_SHAPE;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int W = int(width), nq = int(fri->num_queries);
    const size_t proof_elems = proof_words / 4;
    Scratch S(ctx);
    PermCfgDev cfg_dev;
    void* cfg_blob = nullptr;
    LSP_TRY(upload_air_cfgs(ctx, lookups, n_lookups, cfgs, n_cfgs, width, &cfg_dev, &cfg_blob));
    S.ptrs.push_back(cfg_blob);

    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    struct EvGuard {
        cudaEvent_t a, b;
        ~EvGuard() {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    } ev_guard{ev0, ev1};
    cudaEventRecord(ev0, ctx->stream);

    Fr *proof = nullptr, *sc = nullptr, *betas = nullptr, *evbuf = nullptr;
    LSP_TRY(S.get((void**)&proof, proof_elems * 32));
    LSP_TRY(S.get((void**)&sc, (VT_COUNT + 2) * 32));
    LSP_TRY(S.get((void**)&betas, size_t(n_rounds ? n_rounds : 1) * 32));
    LSP_TRY(S.get((void**)&evbuf, size_t(nq) * (n_rounds ? n_rounds : 1) * 2 * 32));
    const size_t st_per = size_t(ST_ROUND + n_rounds);
    const size_t n_status = 3 + size_t(nq) * st_per;  // [pow_low, ood_bad, non-canonical element, per-query blocks]
    if ((n_status + 1) * 4 > ctx->pinned_bytes) return set_err(ctx, LSP_ERR_PARAM, "too many queries x rounds for the status block");
    int* status = nullptr;
    uint32_t* idx = nullptr;
    DevChallenger* ch = nullptr;
    PathJob* jobs = nullptr;
    const int n_jobs = nq * (2 + n_rounds);
    LSP_TRY(S.get((void**)&status, (n_status + 1) * 4));
    LSP_TRY(S.get((void**)&idx, size_t(nq) * 4));
    LSP_TRY(S.get((void**)&ch, sizeof(DevChallenger)));
    LSP_TRY(S.get((void**)&jobs, size_t(n_jobs) * sizeof(PathJob)));
    LSP_CUDA(ctx, cudaMemcpyAsync(proof, proof_words_in, proof_elems * 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemcpyAsync(sc + VT_COUNT, publics, 64, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemsetAsync(status, 0xff, (n_status + 1) * 4, ctx->stream));  // anything left unwritten reads as a failure

    VerifyArgs A;
    A.proof = proof;
    A.p_local = proof + 2;
    A.p_next = A.p_local + W;
    A.p_chunks = A.p_next + W;
    A.p_commits = A.p_chunks + q;
    A.p_final = A.p_commits + n_rounds;
    const Fr* p_pow = A.p_final + (size_t(1) << log_f);
    A.p_queries = p_pow + 1;
    A.publics = sc + VT_COUNT;
    A.scal = sc;
    A.betas = betas;
    A.idx = idx;
    A.fc = ctx->fc;
    A.W = W;
    A.q = q;
    A.log_n = int(log_n);
    A.log_q = log_q;
    A.log_l = log_l;
    A.n_rounds = n_rounds;
    A.n_final = 1 << log_f;
    A.n_queries = nq;
    A.per_query = (proof_elems - size_t(A.p_queries - proof)) / size_t(nq);
    A.ev = evbuf;
    A.status = status + 3;
    A.ood_bad = status + 1;
    A.cfg = cfg_dev;

    ctx->phase = "verify";
    VerifyTranscriptArgs T;
    T.log_n = int(log_n);
    T.n_rounds = n_rounds;
    T.n_final = A.n_final;
    T.pow_bits = int(fri->proof_of_work_bits);
    T.log_l = log_l;
    T.n_queries = nq;
    T.trace_commit = proof;
    T.quot_commit = proof + 1;
    T.publics = A.publics;
    T.alpha_before_openings = ctx->alpha_before_openings ? 1 : 0;
    T.observe_opened_values = ctx->observe_opened_values ? 1 : 0;
    T.opened = A.p_local;
    T.n_opened = 2 * W + q;
    T.fri_commits = A.p_commits;
    T.final_poly = A.p_final;
    T.pow_witness = p_pow;
    T.scal = sc;
    T.betas = betas;
    T.pow_low = reinterpret_cast<uint32_t*>(status);
    T.idx = idx;
    LSP_CUDA(ctx, cudaMemsetAsync(status + 2, 0, 4, ctx->stream));
    LSP_TRY(verify_transcript(ctx, ch, T));
    LSP_LAUNCH(ctx, k_adopt_sampled_indices, unsigned((nq + 63) / 64), 64, 0, const_cast<Fr*>(A.p_queries), A.per_query, (const uint32_t*)idx, nq);
    LSP_LAUNCH(ctx, k_verify_canonical, grid_for(ctx, proof_elems, 128), 128, 0, (const Fr*)proof, proof_elems, status + 2);
    LSP_LAUNCH(ctx, k_verify_fold_ood, unsigned((nq + VERIFY_BLOCK - 1) / VERIFY_BLOCK + 1), VERIFY_BLOCK, 0, A);
    LSP_LAUNCH(ctx, k_verify_path_jobs, unsigned((n_jobs + 127) / 128), 128, 0, A, jobs);
    LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_merkle_paths_tri<D>, unsigned((n_jobs + 9) / 10), 32, 0, ctx->p2, (const PathJob*)jobs, n_jobs));
    LSP_CUDA(ctx, cudaMemcpyAsync(status + n_status, &ch->overflow, 4, cudaMemcpyDeviceToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, status, (n_status + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(ev1, ctx->stream);
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->phase = "";
    if (device_ms_out) cudaEventElapsedTime(device_ms_out, ev0, ev1);
    const int* h = static_cast<const int*>(ctx->pinned);
    if (h[n_status]) return set_err(ctx, LSP_ERR_STATE, "challenger input buffer overflow");
    // first failing check, in the order the reference verifier meets them
    if (h[2]) return LSP_VERIFY_INVALID_PROOF_SHAPE;   // an element >= r: the proof would not have deserialised
    if (h[0] != 0) return LSP_VERIFY_INVALID_POW_WITNESS;
    for (int qi = 0; qi < nq; qi++) {
        const int* st = h + 3 + size_t(qi) * st_per;
        if (st[ST_INDEX]) return LSP_VERIFY_INVALID_PROOF_SHAPE;
        if (st[ST_TRACE]) return LSP_VERIFY_TRACE_OPENING;
        if (st[ST_QUOT]) return LSP_VERIFY_QUOTIENT_OPENING;
        for (int r = 0; r < n_rounds; r++)
            if (st[ST_ROUND + r]) return LSP_VERIFY_COMMIT_PHASE_OPENING;
        if (st[ST_FINAL]) return LSP_VERIFY_FINAL_POLY_MISMATCH;
    }
    if (h[1]) return LSP_VERIFY_OOD_EVALUATION_MISMATCH;
    return LSP_OK;
}

extern "C" int lsp_verify_permutation(lsp_ctx* ctx, const lsp_fri_config* fri, uint32_t log_n, size_t width, const lsp_perm_air_cfg* cfgs,
                                      int n_cfgs, const uint64_t publics[2][4], const uint64_t* proof_words_in, size_t proof_words,
                                      float* device_ms_out) {
    if (!cfgs || n_cfgs <= 0) return LSP_ERR_PARAM;
    return lsp_verify_air(ctx, fri, log_n, width, nullptr, 0, cfgs, n_cfgs, publics, proof_words_in, proof_words, device_ms_out);
}

// `Mmcs::verify_batch(&commit, &[Dimensions], index, &opened_values, &proof)` for matrices of one height
// (the only case on the reference's path): `row` = the opened rows back to back, as lsp_merkle_open_batch returns them.
extern "C" int lsp_merkle_verify_batch(lsp_ctx* ctx, const uint64_t root[4], uint32_t log_height, size_t index, const uint64_t* row,
                                       size_t row_len, const uint64_t* siblings) {
    if (!ctx || !root || !row || row_len == 0 || row_len > (1u << 20) || log_height > 31 || (log_height && !siblings)) return LSP_ERR_PARAM;
    if (index >> log_height) return LSP_ERR_PARAM;
    if (!ctx->p2_set) return set_err(ctx, LSP_ERR_STATE, "lsp_set_poseidon2 has not been called");
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch S(ctx);
    Fr* buf = nullptr;
    const size_t n_elems = 1 + row_len + log_height;
    LSP_TRY(S.get((void**)&buf, n_elems * 32));
    PathJob* job = nullptr;
    LSP_TRY(S.get((void**)&job, sizeof(PathJob)));
    int* bad = nullptr;
    LSP_TRY(S.get((void**)&bad, 4));
    LSP_CUDA(ctx, cudaMemcpyAsync(buf, root, 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemcpyAsync(buf + 1, row, row_len * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (log_height) LSP_CUDA(ctx, cudaMemcpyAsync(buf + 1 + row_len, siblings, size_t(log_height) * 32, cudaMemcpyHostToDevice, ctx->stream));
    PathJob J;
    J.row = buf + 1;
    J.sibs = buf + 1 + row_len;
    J.root = buf;
    J.bad = bad;
    J.index = uint32_t(index);
    J.width = int(row_len);
    J.log_h = int(log_height);
    LSP_CUDA(ctx, cudaMemcpyAsync(job, &J, sizeof(J), cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemsetAsync(bad, 0xff, 4, ctx->stream));
    LSP_DISPATCH_SBOX(ctx->p2.sbox_d, LSP_LAUNCH(ctx, k_merkle_paths_tri<D>, 1, 32, 0, ctx->p2, (const PathJob*)job, 1));
    LSP_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return *static_cast<const int*>(ctx->pinned) ? LSP_VERIFY_ROOT_MISMATCH : LSP_OK;
}
