// Host driver: `p3_uni_stark::prove` over `TwoAdicFriPcs`, restated as one stream
// of kernel launches with the transcript kept on the device.
//
// Mirrors, stage for stage (span names from the reference's bench.log:18-67):
//   prove(&config, &air, &mut challenger, trace, &publics)       bin/src/main.rs:80-86
//     "commit to trace data"            pcs.commit -> coset_lde_batch + MerkleTreeMmcs::commit
//     "compute quotient polynomial"     quotient_values with LineaAIR::eval (air/src/lib.rs:116-167)
//     "commit to quotient poly chunks"
//     "open"                            opened values, reduced openings, FRI prover
// The host never reads a challenge: alpha, zeta, the FRI betas and the query
// indices are sampled by single-thread kernels from a device-resident
// HashChallenger and consumed by the next kernels straight from HBM, so the
// whole proof is produced without a host<->device round trip; one D2H copy
// returns it.
//
// Proof layout (flat array of Fr, Montgomery limbs; query indices are raw integers):
//   [0] trace commitment  [1] quotient commitment
//   W trace_local, W trace_next, q quotient-chunk values
//   R commit-phase commitments (R = log2 N - log_final_poly_len)
//   F final-poly coefficients (F = 2^(log_blowup + log_final_poly_len))
//   1 pow witness
//   per query: 1 index | W trace row, log2 L siblings | q chunk row, log2 L siblings |
//              per round r: sibling value, (log2 L - 1 - r) siblings
#include "../csrc/stark.cuh"
#include "prove_kernels.cuh"

using namespace lsp;

namespace {

struct FriRoundDev {
    const Fr* folded;   // input vector of this round (len elements)
    const Fr* digests;  // layers over pairs: h = len/2 leaves
    uint32_t log_h;
};

struct QueryArgs {
    const uint32_t* idx;
    const Fr* trace_lde;  size_t lde_rows;  int width;
    const Fr* trace_digests;
    const Fr* const* quot_cols;  int q;
    const Fr* quot_digests;
    int log_l;
    const FriRoundDev* rounds;  int n_rounds;
    Fr* out;  size_t per_query;
};

__device__ __forceinline__ const Fr* tree_sibling(const Fr* digests, size_t h, int k, size_t index) {
    return digests + (2 * h - ((2 * h) >> k)) + ((index >> k) ^ 1);
}

// One block per query; threads stride over the elements to copy.
__global__ void __launch_bounds__(128) k_query_gather(const __grid_constant__ QueryArgs A) {
    const uint32_t index = A.idx[blockIdx.x];
    Fr* out = A.out + size_t(blockIdx.x) * A.per_query;
    const size_t big = size_t(1) << A.log_l;
    if (threadIdx.x == 0) {
        Fr v = fr_zero();
        v.l[0] = index;
        fr_store(out, v);
    }
    size_t o = 1;
    for (int c = threadIdx.x; c < A.width; c += blockDim.x) fr_store(out + o + c, fr_load(A.trace_lde + size_t(c) * A.lde_rows + index));
    o += A.width;
    for (int k = threadIdx.x; k < A.log_l; k += blockDim.x) fr_store(out + o + k, fr_load(tree_sibling(A.trace_digests, big, k, index)));
    o += A.log_l;
    for (int c = threadIdx.x; c < A.q; c += blockDim.x) fr_store(out + o + c, fr_load(A.quot_cols[c] + index));
    o += A.q;
    for (int k = threadIdx.x; k < A.log_l; k += blockDim.x) fr_store(out + o + k, fr_load(tree_sibling(A.quot_digests, big, k, index)));
    o += A.log_l;
    for (int r = 0; r < A.n_rounds; r++) {  // answer_query
        const FriRoundDev R = A.rounds[r];
        const size_t index_i = index >> r;
        if (threadIdx.x == 0) fr_store(out + o, fr_load(R.folded + (index_i ^ 1)));
        o += 1;
        const size_t h = size_t(1) << R.log_h;
        for (int k = threadIdx.x; k < int(R.log_h); k += blockDim.x) fr_store(out + o + k, fr_load(tree_sibling(R.digests, h, k, index_i >> 1)));
        o += R.log_h;
    }
}


// zeta' = zeta * w_N ; chunk points z_c = zeta / (g * w_{Nq}^c)
__global__ void k_open_points(const Fr* __restrict__ zeta, int log_n, int log_q, Fr* __restrict__ zeta_next, Fr* __restrict__ chunk_pts) {
    int c = threadIdx.x;
    Fr z = fr_load(zeta);
    if (c == 0) fr_store(zeta_next, fr_mul(z, fr_two_adic_generator(log_n)));
    if (c < (1 << log_q)) {
        int lnq = log_n + log_q;
        Fr w = fr_two_adic_generator(lnq);
        uint32_t e = uint32_t(((size_t(1) << lnq) - size_t(c)) & ((size_t(1) << lnq) - 1));  // w^-c
        fr_store(chunk_pts + c, fr_mul(fr_mul(z, fr_const(FR_GEN_INV)), fr_pow_u32(w, e)));
    }
}

// shift of quotient chunk c for TwoAdicFriPcs::commit: g / (g * w_{Nq}^c) = w_{Nq}^-c
__global__ void k_chunk_shifts(int log_n, int log_q, Fr* __restrict__ shifts) {
    int c = threadIdx.x;
    if (c < (1 << log_q)) {
        int lnq = log_n + log_q;
        uint32_t e = uint32_t(((size_t(1) << lnq) - size_t(c)) & ((size_t(1) << lnq) - 1));
        fr_store(shifts + c, fr_pow_u32(fr_two_adic_generator(lnq), e));
    }
}

struct ReduceArgs {
    const Fr* trace_lde;  size_t rows;  int width;
    const Fr* const* quot_cols;  int q;
    const Fr* alpha;
    const Fr* s;      // k_open_scalars output
    const Fr* e_zeta; const Fr* e_next;  // inverse denominators
    Fr* out;
};

// reduced opening (FRI input) at every LDE row:
//   ro = [ (Rt - Yt) + a^2W (Rq - Yq) ] / (x - zeta) + a^W (Rt - Yt') / (x - zeta')
__global__ void __launch_bounds__(128) k_reduce_openings(const __grid_constant__ ReduceArgs A) {
    const Fr a = fr_load(A.alpha);
    const Fr yt = fr_load(A.s), ytn = fr_load(A.s + 1), yq = fr_load(A.s + 2), aw = fr_load(A.s + 3), a2w = fr_load(A.s + 4);
    for (size_t p = blockIdx.x * size_t(blockDim.x) + threadIdx.x; p < A.rows; p += size_t(gridDim.x) * blockDim.x) {
        Fr rt = fr_load_nc(A.trace_lde + size_t(A.width - 1) * A.rows + p);
        for (int c = A.width - 2; c >= 0; c--) rt = fr_add(fr_mul(rt, a), fr_load_nc(A.trace_lde + size_t(c) * A.rows + p));
        Fr rq = fr_load_nc(A.quot_cols[A.q - 1] + p);
        for (int c = A.q - 2; c >= 0; c--) rq = fr_add(fr_mul(rq, a), fr_load_nc(A.quot_cols[c] + p));
        Fr t0 = fr_add(fr_sub(rt, yt), fr_mul(a2w, fr_sub(rq, yq)));
        Fr t1 = fr_mul(aw, fr_sub(rt, ytn));
        Fr ro = fr_add(fr_mul(t0, fr_load_nc(A.e_zeta + p)), fr_mul(t1, fr_load_nc(A.e_next + p)));
        fr_store(A.out + p, ro);
    }
}


// round r: input vector at folded_all + (2L - 2L>>r), digests at fri_digests + (2L - 2L>>r) - r
__global__ void k_make_rounds(const Fr* folded_all, const Fr* fri_digests, int log_l, int n_rounds, FriRoundDev* out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rounds) return;
    size_t two_l = size_t(2) << log_l;
    size_t off = two_l - (two_l >> r);
    out[r].folded = folded_all + off;
    out[r].digests = fri_digests + off - r;
    out[r].log_h = uint32_t(log_l - 1 - r);
}



int max_constraint_log_quotient(int n_lookups, int /*n_perms*/) {
    // `get_log_quotient_degree`, log2_ceil(max degree - 1):
    //   permutation AIR: max degree 3 (is_first_row * (check - a_ch * inv), air/src/lib.rs:146-148)      => 1
    //   lookup AIR:      max degree 4 (is_first_row * (check - filter * inverse + ...), :100-102)        => 2
    // whatever the column counts.
    return n_lookups > 0 ? 2 : 1;
}

}  // namespace

extern "C" size_t lsp_proof_words(uint32_t log_n, uint32_t width, uint32_t log_q, const lsp_fri_config* fri) {
    if (!fri || fri->log_final_poly_len > log_n) return 0;
    size_t log_l = size_t(log_n) + fri->log_blowup;
    size_t q = size_t(1) << log_q;
    size_t rounds = log_n - fri->log_final_poly_len;
    size_t f = size_t(1) << (fri->log_blowup + fri->log_final_poly_len);
    size_t per_query = 1 + (width + log_l) + (q + log_l);
    for (size_t r = 0; r < rounds; r++) per_query += 1 + (log_l - 1 - r);
    size_t elems = 2 + 2 * size_t(width) + q + rounds + f + 1 + size_t(fri->num_queries) * per_query;
    return elems * 4;
}

extern "C" int lsp_air_log_quotient_degree(int n_lookups, int n_perms) { return max_constraint_log_quotient(n_lookups, n_perms); }

extern "C" int lsp_prove_air_dev(lsp_ctx* ctx, const lsp_fri_config* fri, const lsp_mat* trace, const lsp_lookup_air_cfg* lookups,
                                 int n_lookups, const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                                 uint64_t* proof_out, size_t proof_words, float* timings_ms_out) {
    if (!ctx || !fri || !trace || !publics || !proof_out || n_lookups < 0 || n_cfgs < 0 || n_lookups + n_cfgs <= 0) return LSP_ERR_PARAM;
    if ((n_cfgs && !cfgs) || (n_lookups && !lookups)) return LSP_ERR_PARAM;
    if (!ctx->p2_set) return set_err(ctx, LSP_ERR_STATE, "lsp_set_poseidon2 has not been called");
    const size_t n = trace->rows, W = trace->width;
    if (!is_pow2(n)) return set_err(ctx, LSP_ERR_PARAM, "trace height %zu is not a power of two (prove would panic)", n);
    const int log_n = ilog2(n);
    const int log_q = max_constraint_log_quotient(n_lookups, n_cfgs);
    const int q = 1 << log_q;
    const int log_b = int(fri->log_blowup), log_l = log_n + log_b;
    if (log_q > log_b) return set_err(ctx, LSP_ERR_PARAM, "quotient degree 2^%d exceeds blowup 2^%d", log_q, log_b);
    if (int(fri->log_final_poly_len) > log_n) return set_err(ctx, LSP_ERR_PARAM, "log_final_poly_len exceeds log2 of the trace height");
    if (log_l > 31 || log_l < 1) return set_err(ctx, LSP_ERR_PARAM, "LDE of 2^%d rows unsupported", log_l);
    if (fri->num_queries == 0 || fri->num_queries > 4096) return set_err(ctx, LSP_ERR_PARAM, "num_queries out of range");
    const size_t need = lsp_proof_words(log_n, uint32_t(W), log_q, fri);
    if (proof_words < need) return set_err(ctx, LSP_ERR_PARAM, "proof buffer too small: %zu < %zu words", proof_words, need);
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));

    const size_t L = size_t(1) << log_l;
    const int n_rounds = log_n - int(fri->log_final_poly_len);
    const int log_f = log_b + int(fri->log_final_poly_len);
    const size_t proof_elems = need / 4;
    Scratch S(ctx);

    cudaEvent_t ev[9];
    for (auto& e : ev) cudaEventCreate(&e);
    int n_ev = 0;
    auto mark = [&]() { cudaEventRecord(ev[n_ev++], ctx->stream); };
    struct EvGuard {
        cudaEvent_t* e;
        ~EvGuard() {
            for (int i = 0; i < 9; i++) cudaEventDestroy(e[i]);
        }
    } ev_guard{ev};

    // ---- device-side scalars and transcript ---------------------------------
    enum { S_PUB0, S_PUB1, S_LOGN, S_ALPHA, S_ZETA, S_ZETA_NEXT, S_ALPHA_FRI, S_GEN, S_CHUNK_SHIFT, S_CHUNK_PT = S_CHUNK_SHIFT + 8,
           S_OPEN = S_CHUNK_PT + 8, S_COUNT = S_OPEN + 8 };
    Fr* sc = nullptr;
    LSP_TRY(S.get((void**)&sc, S_COUNT * 32));
    LSP_CUDA(ctx, cudaMemcpyAsync(sc + S_PUB0, publics, 64, cudaMemcpyHostToDevice, ctx->stream));
    LSP_LAUNCH(ctx, k_set_small, 1, 1, 0, sc + S_LOGN, uint32_t(log_n));
    LSP_LAUNCH(ctx, k_set_small, 1, 1, 0, sc + S_GEN, 22u);
    DevChallenger* ch = nullptr;
    LSP_TRY(S.get((void**)&ch, sizeof(DevChallenger)));
    LSP_TRY(challenger_init(ctx, ch));
    Fr* proof = nullptr;
    LSP_TRY(S.get((void**)&proof, proof_elems * 32));
    Fr* p_trace_commit = proof;
    Fr* p_quot_commit = proof + 1;
    Fr* p_local = proof + 2;
    Fr* p_next = p_local + W;
    Fr* p_chunks = p_next + W;
    Fr* p_fri_commits = p_chunks + q;
    Fr* p_final = p_fri_commits + n_rounds;
    Fr* p_pow = p_final + (size_t(1) << log_f);
    Fr* p_queries = p_pow + 1;
    PermCfgDev cfg_dev;
    void* cfg_blob = nullptr;
    LSP_TRY(upload_air_cfgs(ctx, lookups, n_lookups, cfgs, n_cfgs, W, &cfg_dev, &cfg_blob));  // also checks the AIR width
    S.ptrs.push_back(cfg_blob);
    mark();  // 0

    // ---- commit to trace data ---------------------------------------------------
    ctx->phase = "commit_trace";
    Fr *coef_t = nullptr, *lde_t = nullptr, *dig_t = nullptr;
    LSP_TRY(S.get((void**)&coef_t, n * W * 32));
    LSP_TRY(S.get((void**)&lde_t, L * W * 32));
    LSP_TRY(S.get((void**)&dig_t, (2 * L - 1) * 32));
    LSP_TRY(interpolate_columns(ctx, trace->d, n, W, coef_t));
    LSP_TRY(coset_evaluate(ctx, coef_t, n, W, log_b, sc + S_GEN, lde_t));  // shift = GENERATOR / 1
    mark();  // 1
    const Fr** cols_t = nullptr;
    LSP_TRY(S.get((void**)&cols_t, W * sizeof(Fr*)));
    LSP_LAUNCH(ctx, k_make_cols, unsigned((W + 63) / 64), 64, 0, (const Fr*)lde_t, L, int(W), cols_t);
    LSP_TRY(merkle_build(ctx, cols_t, int(W), L, dig_t));
    LSP_CUDA(ctx, cudaMemcpyAsync(p_trace_commit, dig_t + (2 * L - 2), 32, cudaMemcpyDeviceToDevice, ctx->stream));
    mark();  // 2

    // observe(log_degree); observe(trace_commit); observe_slice(publics); alpha <- sample
    LSP_TRY(challenger_observe_dev(ctx, ch, sc + S_LOGN, 1));
    LSP_TRY(challenger_observe_dev(ctx, ch, p_trace_commit, 1));
    LSP_TRY(challenger_observe_dev(ctx, ch, sc + S_PUB0, 2));
    LSP_TRY(challenger_sample(ctx, ch, sc + S_ALPHA));

    // ---- compute quotient polynomial -------------------------------------------
    ctx->phase = "quotient";
    Fr* chunks = nullptr;
    LSP_TRY(S.get((void**)&chunks, size_t(q) * n * 32));
    LSP_TRY(quotient_permutation(ctx, lde_t, L, log_n, log_q, cfg_dev, sc + S_PUB0, sc + S_ALPHA, chunks));
    mark();  // 3

    // ---- commit to quotient poly chunks ----------------------------------------
    ctx->phase = "commit_quotient";
    Fr *coef_q = nullptr, *lde_q = nullptr, *dig_q = nullptr;
    LSP_TRY(S.get((void**)&coef_q, size_t(q) * n * 32));
    LSP_TRY(S.get((void**)&lde_q, size_t(q) * L * 32));
    LSP_TRY(S.get((void**)&dig_q, (2 * L - 1) * 32));
    LSP_LAUNCH(ctx, k_chunk_shifts, 1, 32, 0, log_n, log_q, sc + S_CHUNK_SHIFT);
    LSP_TRY(interpolate_columns(ctx, chunks, n, q, coef_q));
    for (int c = 0; c < q; c++)
        LSP_TRY(coset_evaluate(ctx, coef_q + size_t(c) * n, n, 1, log_b, sc + S_CHUNK_SHIFT + c, lde_q + size_t(c) * L));
    const Fr** cols_q = nullptr;
    LSP_TRY(S.get((void**)&cols_q, q * sizeof(Fr*)));
    LSP_LAUNCH(ctx, k_make_cols, 1, 64, 0, (const Fr*)lde_q, L, q, cols_q);
    LSP_TRY(merkle_build(ctx, cols_q, q, L, dig_q));
    LSP_CUDA(ctx, cudaMemcpyAsync(p_quot_commit, dig_q + (2 * L - 2), 32, cudaMemcpyDeviceToDevice, ctx->stream));
    mark();  // 4

    // observe(quotient_commit); zeta <- sample; zeta_next = zeta * w_N
    LSP_TRY(challenger_observe_dev(ctx, ch, p_quot_commit, 1));
    LSP_TRY(challenger_sample(ctx, ch, sc + S_ZETA));
    LSP_LAUNCH(ctx, k_open_points, 1, 32, 0, sc + S_ZETA, log_n, log_q, sc + S_ZETA_NEXT, sc + S_CHUNK_PT);

    // ---- open --------------------------------------------------------------------
    ctx->phase = "open";
    // (fork-era order) the batching challenge is sampled before the openings
    LSP_TRY(challenger_sample(ctx, ch, sc + S_ALPHA_FRI));
    LSP_TRY(eval_columns_at(ctx, coef_t, n, W, sc + S_ZETA, p_local));
    LSP_TRY(eval_columns_at(ctx, coef_t, n, W, sc + S_ZETA_NEXT, p_next));
    for (int c = 0; c < q; c++)  // chunk c's interpolant lives on the shifted domain: evaluate at zeta/(g w^c)
        LSP_TRY(eval_columns_at(ctx, coef_q + size_t(c) * n, n, 1, sc + S_CHUNK_PT + c, p_chunks + c));
    LSP_LAUNCH(ctx, k_open_scalars, 1, 1, 0, sc + S_ALPHA_FRI, p_local, p_next, p_chunks, int(W), q, sc + S_OPEN);
    Fr* inv_den[2] = {nullptr, nullptr};
    LSP_TRY(S.get((void**)&inv_den[0], L * 32));
    LSP_TRY(S.get((void**)&inv_den[1], L * 32));
    LSP_TRY(inverse_denominators(ctx, sc + S_ZETA, 2, log_l, inv_den));  // S_ZETA, S_ZETA_NEXT are adjacent
    // folded vectors of all rounds live in one buffer: L + L/2 + ... < 2L
    Fr* folded_all = nullptr;
    LSP_TRY(S.get((void**)&folded_all, 2 * L * 32));
    {
        ReduceArgs A;
        A.trace_lde = lde_t;
        A.rows = L;
        A.width = int(W);
        A.quot_cols = cols_q;
        A.q = q;
        A.alpha = sc + S_ALPHA_FRI;
        A.s = sc + S_OPEN;
        A.e_zeta = inv_den[0];
        A.e_next = inv_den[1];
        A.out = folded_all;
        LSP_LAUNCH(ctx, k_reduce_openings, grid_for(ctx, L, 128), 128, 0, A);
    }
    mark();  // 5

    // ---- FRI commit phase ------------------------------------------------------
    ctx->phase = "fri_commit";
    Fr* fri_digests = nullptr;
    LSP_TRY(S.get((void**)&fri_digests, 2 * L * 32));
    Fr* beta = nullptr;
    LSP_TRY(S.get((void**)&beta, 32));
    {
        Fr* cur = folded_all;
        Fr* dig = fri_digests;
        size_t len = L;
        for (int r = 0; r < n_rounds; r++) {
            LSP_TRY(merkle_build_pairs(ctx, cur, len, dig));
            const Fr* root = dig + (len - 2);
            LSP_CUDA(ctx, cudaMemcpyAsync(p_fri_commits + r, root, 32, cudaMemcpyDeviceToDevice, ctx->stream));
            LSP_TRY(challenger_observe_dev(ctx, ch, root, 1));
            LSP_TRY(challenger_sample(ctx, ch, beta));
            Fr* nxt = cur + len;
            LSP_TRY(fri_fold(ctx, cur, len, beta, nxt));
            dig += len - 1;
            cur = nxt;
            len >>= 1;
        }
        // final polynomial: bit-reverse, iDFT, observe every coefficient
        LSP_LAUNCH(ctx, k_final_poly, 1, unsigned(len < 32 ? 32 : len), 0, (const Fr*)cur, log_f, host_pow2_inverse(log_f), p_final);
        LSP_TRY(challenger_observe_dev(ctx, ch, p_final, int(len)));
    }
    mark();  // 6

    // ---- grind + query phase ---------------------------------------------------
    ctx->phase = "fri_query";
    LSP_TRY(challenger_grind(ctx, ch, int(fri->proof_of_work_bits), p_pow));
    uint32_t* idx = nullptr;
    LSP_TRY(S.get((void**)&idx, fri->num_queries * 4));
    LSP_TRY(challenger_sample_bits(ctx, ch, log_l, int(fri->num_queries), idx));
    FriRoundDev* rounds_dev = nullptr;
    LSP_TRY(S.get((void**)&rounds_dev, (n_rounds ? n_rounds : 1) * sizeof(FriRoundDev)));
    if (n_rounds)
        LSP_LAUNCH(ctx, k_make_rounds, unsigned((n_rounds + 63) / 64), 64, 0, (const Fr*)folded_all, (const Fr*)fri_digests, log_l, n_rounds, rounds_dev);
    {
        QueryArgs A;
        A.idx = idx;
        A.trace_lde = lde_t;
        A.lde_rows = L;
        A.width = int(W);
        A.trace_digests = dig_t;
        A.quot_cols = cols_q;
        A.q = q;
        A.quot_digests = dig_q;
        A.log_l = log_l;
        A.rounds = rounds_dev;
        A.n_rounds = n_rounds;
        A.out = p_queries;
        A.per_query = (proof_elems - size_t(p_queries - proof)) / fri->num_queries;
        LSP_LAUNCH(ctx, k_query_gather, fri->num_queries, 128, 0, A);
    }
    mark();  // 7
    LSP_CUDA(ctx, cudaMemcpyAsync(proof_out, proof, proof_elems * 32, cudaMemcpyDeviceToHost, ctx->stream));
    mark();  // 8
    ctx->phase = "";
    LSP_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, &ch->overflow, 4, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int overflow = *(volatile int*)ctx->pinned;
    if (overflow) return set_err(ctx, LSP_ERR_STATE, "challenger input buffer overflow");
    if (timings_ms_out)
        for (int i = 0; i < 8; i++) cudaEventElapsedTime(&timings_ms_out[i], ev[i], ev[i + 1]);
    return LSP_OK;
}

extern "C" int lsp_prove_air(lsp_ctx* ctx, const lsp_fri_config* fri, const uint64_t* trace, size_t rows, size_t width,
                             const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* cfgs, int n_cfgs,
                             const uint64_t publics[2][4], uint64_t* proof_out, size_t proof_words, float* timings_ms_out) {
    if (!ctx || !trace) return LSP_ERR_PARAM;
    lsp_mat* m = nullptr;
    LSP_TRY(lsp_mat_upload(ctx, trace, rows, width, &m));
    int rc = lsp_prove_air_dev(ctx, fri, m, lookups, n_lookups, cfgs, n_cfgs, publics, proof_out, proof_words, timings_ms_out);
    lsp_mat_free(ctx, m);
    return rc;
}

extern "C" int lsp_prove_permutation_dev(lsp_ctx* ctx, const lsp_fri_config* fri, const lsp_mat* trace, const lsp_perm_air_cfg* cfgs,
                                         int n_cfgs, const uint64_t publics[2][4], uint64_t* proof_out, size_t proof_words,
                                         float* timings_ms_out) {
    if (!cfgs || n_cfgs <= 0) return LSP_ERR_PARAM;
    return lsp_prove_air_dev(ctx, fri, trace, nullptr, 0, cfgs, n_cfgs, publics, proof_out, proof_words, timings_ms_out);
}

extern "C" int lsp_prove_permutation(lsp_ctx* ctx, const lsp_fri_config* fri, const uint64_t* trace, size_t rows, size_t width,
                                     const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4], uint64_t* proof_out,
                                     size_t proof_words, float* timings_ms_out) {
    if (!cfgs || n_cfgs <= 0) return LSP_ERR_PARAM;
    return lsp_prove_air(ctx, fri, trace, rows, width, nullptr, 0, cfgs, n_cfgs, publics, proof_out, proof_words, timings_ms_out);
}

// ---------------------------------------------------------------------------
// standalone entry points for the pieces (parity tests / trait-level drop-in)
// ---------------------------------------------------------------------------
extern "C" int lsp_quotient_permutation(lsp_ctx* ctx, const lsp_mat* lde_bitrev, int log_n, int log_q, const lsp_perm_air_cfg* cfgs,
                                        int n_cfgs, const uint64_t publics[2][4], const uint64_t alpha[4], lsp_mat** chunks_out) {
    if (!cfgs || n_cfgs <= 0) return LSP_ERR_PARAM;
    return lsp_quotient_air(ctx, lde_bitrev, log_n, log_q, nullptr, 0, cfgs, n_cfgs, publics, alpha, chunks_out);
}

extern "C" int lsp_quotient_air(lsp_ctx* ctx, const lsp_mat* lde_bitrev, int log_n, int log_q, const lsp_lookup_air_cfg* lookups,
                                int n_lookups, const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                                const uint64_t alpha[4], lsp_mat** chunks_out) {
    if (!ctx || !lde_bitrev || !publics || !alpha || !chunks_out || log_n < 0 || log_q < 0 || log_n + log_q > 31) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch S(ctx);
    PermCfgDev cfg_dev;
    void* blob = nullptr;
    LSP_TRY(upload_air_cfgs(ctx, lookups, n_lookups, cfgs, n_cfgs, lde_bitrev->width, &cfg_dev, &blob));
    S.ptrs.push_back(blob);
    Fr* sc = nullptr;
    LSP_TRY(S.get((void**)&sc, 3 * 32));
    LSP_CUDA(ctx, cudaMemcpyAsync(sc, publics, 64, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaMemcpyAsync(sc + 2, alpha, 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    lsp_mat* out = nullptr;
    LSP_TRY(mat_alloc(ctx, size_t(1) << log_n, size_t(1) << log_q, &out));
    int rc = quotient_permutation(ctx, lde_bitrev->d, lde_bitrev->rows, log_n, log_q, cfg_dev, sc, sc + 2, out->d);
    if (rc != LSP_OK) {
        lsp_mat_free(ctx, out);
        return rc;
    }
    *chunks_out = out;
    return LSP_OK;
}

extern "C" int lsp_fri_fold(lsp_ctx* ctx, const lsp_mat* in, const uint64_t beta[4], lsp_mat** out) {
    if (!ctx || !in || !beta || !out || in->width != 1 || !is_pow2(in->rows) || in->rows < 2) return LSP_ERR_PARAM;
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch S(ctx);
    Fr* b = nullptr;
    LSP_TRY(S.get((void**)&b, 32));
    LSP_CUDA(ctx, cudaMemcpyAsync(b, beta, 32, cudaMemcpyHostToDevice, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    lsp_mat* o = nullptr;
    LSP_TRY(mat_alloc(ctx, in->rows / 2, 1, &o));
    int rc = fri_fold(ctx, in->d, in->rows, b, o->d);
    if (rc != LSP_OK) {
        lsp_mat_free(ctx, o);
        return rc;
    }
    *out = o;
    return LSP_OK;
}
