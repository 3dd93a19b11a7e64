// `p3_uni_stark::prove` over `TwoAdicFriPcs` (bin/src/main.rs:80-86), THE implementation: one proof on G >= 1
// GPUs, sharded by ROW RANGES of the bit-reversed LDE.  The single-GPU entry points
// (lsp_prove_air[_dev], host/prover.cu) call this code with a one-rank communicator, so there is one prover to keep
// bit-exact, not two.  Rank r owns storage rows [r*L/G, (r+1)*L/G) of every committed matrix,
// i.e. 2^log_blowup / G whole cosets of the evaluation domain, so
//   * the coset LDE of its rows needs no traffic (coefficients are replicated: every rank
//     interpolates the full trace once);
//   * leaf hashing and the bottom log2(L/G) Merkle levels are local; an all-gather of G
//     subtree roots (32 B each) lets every rank finish the top log2 G levels itself;
//   * the quotient chunk c lives entirely in row block bitrev(c): its owner computes it and
//     broadcasts N values (the one bulk exchange of the proof);
//   * FRI folds pair adjacent rows, so commit-phase rounds stay local (+ one root all-gather)
//     until the layer is small, then the vector is all-gathered once and finished everywhere;
//   * the transcript is replicated: every rank runs the same device challenger on the same
//     roots, so challenges never travel;
//   * query openings are written by their owners into a zeroed buffer and combined with one
//     integer all-reduce(sum).
// The proof is bit-identical to the single-GPU one (tests/test_gpu_sharded.py).
//
// Collectives go through NCCL (dlopen'ed: the library has no link-time dependency on it).
// `lsp_comm_init_local` instead hosts all G ranks in ONE process on ONE device, running the
// ranks' stages in lockstep and replacing every collective by device copies: the same code
// path, testable on a single GPU (and what the 1-GPU CI box exercises).
#include <algorithm>
#include <dlfcn.h>
#include <nccl.h>

#include "../csrc/stark.cuh"
#include "comm.hpp"
#include "prove_kernels.cuh"

using namespace lsp;

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* load_nccl() {
    static NcclApi api;
    // RTLD_NOLOAD first: reuse the copy torch.distributed already mapped, if any
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (!h) return nullptr;
    api.lib = h;
    *(void**)&api.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void**)&api.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void**)&api.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void**)&api.AllGather = dlsym(h, "ncclAllGather");
    *(void**)&api.Broadcast = dlsym(h, "ncclBroadcast");
    *(void**)&api.AllReduce = dlsym(h, "ncclAllReduce");
    *(void**)&api.GetErrorString = dlsym(h, "ncclGetErrorString");
    *(void**)&api.Send = dlsym(h, "ncclSend");
    *(void**)&api.Recv = dlsym(h, "ncclRecv");
    *(void**)&api.GroupStart = dlsym(h, "ncclGroupStart");
    *(void**)&api.GroupEnd = dlsym(h, "ncclGroupEnd");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.Broadcast || !api.AllReduce || !api.Send ||
        !api.Recv || !api.GroupStart || !api.GroupEnd) {
        api.lib = nullptr;
        return nullptr;
    }
    return &api;
}

// Loaded once; a function-local static is initialised under a lock, so the host threads of a
// one-process-many-GPUs driver (lsp_prove --gpus N) may all arrive here at once.
NcclApi* nccl_api() {
    static NcclApi* const api = load_nccl();
    return api;
}

}  // namespace

namespace {

#define LSP_NCCL(ctx, call)                                                                                   \
    do {                                                                                                      \
        ncclResult_t r__ = (call);                                                                            \
        if (r__ != ncclSuccess)                                                                               \
            return set_err(ctx, LSP_ERR_COMM, "%s:%d %s: %s", __FILE__, __LINE__, #call,                     \
                           nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r__) : "nccl error");      \
    } while (0)

// ---- collectives over the ranks hosted by this process --------------------------------------
// send[i] / recv[i] are the buffers of hosted rank i (ranks[i] is its global rank).
int coll_allgather(lsp_comm* cm, const std::vector<int>& ranks, const std::vector<const void*>& send, const std::vector<void*>& recv,
                   size_t bytes) {
    lsp_ctx* ctx = cm->ctx;
    if (cm->local) {
        for (size_t d = 0; d < ranks.size(); d++)
            for (size_t s = 0; s < ranks.size(); s++)
                LSP_CUDA(ctx, cudaMemcpyAsync((char*)recv[d] + size_t(ranks[s]) * bytes, send[s], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return LSP_OK;
    }
    LSP_NCCL(ctx, nccl_api()->AllGather(send[0], recv[0], bytes / 8, ncclUint64, ((ncclComm_t)cm->nccl), ctx->stream));
    return LSP_OK;
}
int coll_broadcast(lsp_comm* cm, const std::vector<int>& ranks, int root, const std::vector<void*>& buf, size_t bytes) {
    lsp_ctx* ctx = cm->ctx;
    if (cm->local) {
        for (size_t d = 0; d < ranks.size(); d++)
            if (ranks[d] != root)
                LSP_CUDA(ctx, cudaMemcpyAsync(buf[d], buf[root], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return LSP_OK;
    }
    LSP_NCCL(ctx, nccl_api()->Broadcast(buf[0], buf[0], bytes / 8, ncclUint64, root, ((ncclComm_t)cm->nccl), ctx->stream));
    return LSP_OK;
}

// Point-to-point: hosted rank i sends send[i] to rank send_to[i] and receives recv[i] from rank recv_from[i] (-1: takes no
// part).  Local mode: a device copy from the sender's buffer (send[] is indexed by global rank there).
int coll_sendrecv(lsp_comm* cm, const std::vector<int>& ranks, const std::vector<int>& send_to, const std::vector<int>& recv_from,
                  const std::vector<const void*>& send, const std::vector<void*>& recv, size_t bytes) {
    lsp_ctx* ctx = cm->ctx;
    if (cm->local) {
        for (size_t i = 0; i < ranks.size(); i++)
            if (recv_from[i] >= 0)
                LSP_CUDA(ctx, cudaMemcpyAsync(recv[i], send[size_t(recv_from[i])], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return LSP_OK;
    }
    if (send_to[0] < 0 && recv_from[0] < 0) return LSP_OK;
    LSP_NCCL(ctx, nccl_api()->GroupStart());
    if (send_to[0] >= 0) LSP_NCCL(ctx, nccl_api()->Send(send[0], bytes / 8, ncclUint64, send_to[0], ((ncclComm_t)cm->nccl), ctx->stream));
    if (recv_from[0] >= 0) LSP_NCCL(ctx, nccl_api()->Recv(recv[0], bytes / 8, ncclUint64, recv_from[0], ((ncclComm_t)cm->nccl), ctx->stream));
    LSP_NCCL(ctx, nccl_api()->GroupEnd());
    return LSP_OK;
}

__global__ void k_sum_u64(unsigned long long* __restrict__ dst, const unsigned long long* const* __restrict__ srcs, int n_src, size_t n) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        unsigned long long a = 0;
        for (int s = 0; s < n_src; s++) a += srcs[s][i];
        dst[i] = a;
    }
}
// In local mode the result lands in `out` (one shared buffer); in NCCL mode in place in buf[0].
int coll_allreduce_sum(lsp_comm* cm, const std::vector<int>& ranks, const std::vector<void*>& buf, size_t bytes, void* local_out) {
    lsp_ctx* ctx = cm->ctx;
    if (cm->local) {
        const unsigned long long** d_srcs = nullptr;
        LSP_TRY(dev_alloc(ctx, (void**)&d_srcs, ranks.size() * sizeof(void*)));
        memcpy(ctx->pinned, buf.data(), ranks.size() * sizeof(void*));
        LSP_CUDA(ctx, cudaMemcpyAsync(d_srcs, ctx->pinned, ranks.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
        LSP_LAUNCH(ctx, k_sum_u64, grid_for(ctx, bytes / 8, 256), 256, 0, (unsigned long long*)local_out, d_srcs, int(ranks.size()), bytes / 8);
        LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // pinned staging is reused by the caller
        dev_free(ctx, (void*)d_srcs);
        return LSP_OK;
    }
    LSP_NCCL(ctx, nccl_api()->AllReduce(buf[0], buf[0], bytes / 8, ncclUint64, ncclSum, ((ncclComm_t)cm->nccl), ctx->stream));
    return LSP_OK;
}

// ---- device helpers --------------------------------------------------------------------------
__global__ void k_chunk_consts(const FieldConsts* __restrict__ fc, const Fr* __restrict__ zeta, int log_n, int log_q, Fr* __restrict__ shifts, Fr* __restrict__ zeta_next,
                               Fr* __restrict__ chunk_pts) {
    int c = threadIdx.x;
    int lnq = log_n + log_q;
    if (c < (1 << log_q)) {
        uint32_t e = uint32_t(((size_t(1) << lnq) - size_t(c)) & ((size_t(1) << lnq) - 1));
        Fr wi = fr_pow_u32(fr_two_adic_generator(fc, lnq), e);  // w_{Nq}^-c
        if (shifts) fr_store(shifts + c, wi);
        if (zeta) fr_store(chunk_pts + c, fr_mul(fr_mul(fr_load(zeta), fr_load(&fc->gen_inv)), wi));
    }
    if (c == 0 && zeta) fr_store(zeta_next, fr_mul(fr_load(zeta), fr_two_adic_generator(fc, log_n)));
}
struct ReduceArgsS {
    const Fr* trace_lde; size_t rows; int width;
    const Fr* quot_lde; int q;   // q columns, stride rows
    const Fr* alpha; const Fr* s; const Fr* e_zeta; const Fr* e_next;
    Fr* out;
};
__global__ void __launch_bounds__(128) k_reduce_openings_s(const __grid_constant__ ReduceArgsS A) {
    const Fr a = fr_load(A.alpha);
    const Fr yt = fr_load(A.s), ytn = fr_load(A.s + 1), yq = fr_load(A.s + 2), aw = fr_load(A.s + 3), a2w = fr_load(A.s + 4);
    for (size_t p = blockIdx.x * size_t(blockDim.x) + threadIdx.x; p < A.rows; p += size_t(gridDim.x) * blockDim.x) {
        Fr rt = fr_load_nc(A.trace_lde + size_t(A.width - 1) * A.rows + p);
        for (int c = A.width - 2; c >= 0; c--) rt = fr_add(fr_mul(rt, a), fr_load_nc(A.trace_lde + size_t(c) * A.rows + p));
        Fr rq = fr_load_nc(A.quot_lde + size_t(A.q - 1) * A.rows + p);
        for (int c = A.q - 2; c >= 0; c--) rq = fr_add(fr_mul(rq, a), fr_load_nc(A.quot_lde + size_t(c) * A.rows + p));
        Fr t0 = fr_add(fr_sub(rt, yt), fr_mul(a2w, fr_sub(rq, yq)));
        Fr t1 = fr_mul(aw, fr_sub(rt, ytn));
        fr_store(A.out + p, fr_add(fr_mul(t0, fr_load_nc(A.e_zeta + p)), fr_mul(t1, fr_load_nc(A.e_next + p))));
    }
}

// A Merkle tree whose leaves are split over G ranks: local layers + replicated top.
struct ShardTree {
    const Fr* local;  // local digest layers (2*hl - 1), hl = h / G leaves
    const Fr* top;    // 2G - 1 digests: layer 0 = the G subtree roots
    uint32_t log_h;   // global height
};
constexpr int LSP_MAX_FRI_ROUNDS = 32;   // log2 of the LDE height <= 31
struct FriRoundS {
    const Fr* vec;    // round input: local slice (sharded) or the whole vector (replicated)
    ShardTree tree;   // for replicated rounds: local = the full tree, top unused
    uint32_t sharded;
};
struct QueryArgsS {
    const uint32_t* idx;
    int rank, log_g;
    const Fr* trace_lde; size_t rows; int width;   // local rows, column stride = rows
    ShardTree trace_tree;
    const Fr* quot_lde; int q;
    ShardTree quot_tree;
    int log_l;
    FriRoundS rounds[LSP_MAX_FRI_ROUNDS]; int n_rounds;   // by value: no staging copy, no host synchronisation
    Fr* out; size_t per_query;
};

// sibling of `index` at level k of a sharded tree; returns false if this rank does not hold it
__device__ __forceinline__ bool shard_sibling(const ShardTree& t, int k, size_t index, int rank, int log_g, Fr& out) {
    const int log_hl = int(t.log_h) - log_g;
    const size_t node = (index >> k) ^ 1;
    if (k < log_hl) {
        const size_t per_rank = size_t(1) << (log_hl - k);
        if (int(node / per_rank) != rank) return false;
        const size_t hl = size_t(1) << log_hl;
        out = fr_load(t.local + (2 * hl - ((2 * hl) >> k)) + (node & (per_rank - 1)));
        return true;
    }
    if (rank != 0) return false;  // replicated part: rank 0 speaks
    const int kt = k - log_hl;
    const size_t g = size_t(1) << log_g;
    out = fr_load(t.top + (2 * g - ((2 * g) >> kt)) + node);
    return true;
}

// One block per query.  Every element of the query section is written by exactly one rank
// (the buffer is zeroed first), so an integer all-reduce(sum) assembles it.
__global__ void __launch_bounds__(128) k_query_gather_s(const __grid_constant__ QueryArgsS A) {
    const uint32_t index = A.idx[blockIdx.x];
    Fr* out = A.out + size_t(blockIdx.x) * A.per_query;
    const int rank = A.rank, log_g = A.log_g;
    if (threadIdx.x == 0 && rank == 0) {
        Fr v = fr_zero();
        v.l[0] = index;
        fr_store(out, v);
    }
    size_t o = 1;
    const bool own_row = int(index / A.rows) == rank;
    const size_t lrow = index % A.rows;
    Fr v;
    if (own_row)
        for (int c = threadIdx.x; c < A.width; c += blockDim.x) fr_store(out + o + c, fr_load(A.trace_lde + size_t(c) * A.rows + lrow));
    o += A.width;
    for (int k = threadIdx.x; k < A.log_l; k += blockDim.x)
        if (shard_sibling(A.trace_tree, k, index, rank, log_g, v)) fr_store(out + o + k, v);
    o += A.log_l;
    if (own_row)
        for (int c = threadIdx.x; c < A.q; c += blockDim.x) fr_store(out + o + c, fr_load(A.quot_lde + size_t(c) * A.rows + lrow));
    o += A.q;
    for (int k = threadIdx.x; k < A.log_l; k += blockDim.x)
        if (shard_sibling(A.quot_tree, k, index, rank, log_g, v)) fr_store(out + o + k, v);
    o += A.log_l;
    for (int r = 0; r < A.n_rounds; r++) {
        const FriRoundS& R = A.rounds[r];
        const size_t index_i = index >> r;
        const int log_h = int(R.tree.log_h);  // pairs
        if (R.sharded) {
            const size_t local_len = (size_t(2) << log_h) >> log_g;
            const size_t sib = index_i ^ 1;
            if (threadIdx.x == 0 && int(sib / local_len) == rank) fr_store(out + o, fr_load(R.vec + (sib % local_len)));
            o += 1;
            for (int k = threadIdx.x; k < log_h; k += blockDim.x)
                if (shard_sibling(R.tree, k, index_i >> 1, rank, log_g, v)) fr_store(out + o + k, v);
        } else {
            if (rank == 0) {
                if (threadIdx.x == 0) fr_store(out + o, fr_load(R.vec + (index_i ^ 1)));
                const size_t h = size_t(1) << log_h;
                for (int k = threadIdx.x; k < log_h; k += blockDim.x)
                    fr_store(out + o + 1 + k, fr_load(R.tree.local + (2 * h - ((2 * h) >> k)) + (((index_i >> 1) >> k) ^ 1)));
            }
            o += 1;
        }
        o += log_h;
    }
}

uint32_t bitrev_host(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

struct Pool {  // frees everything on scope exit
    lsp_ctx* ctx;
    std::vector<void*> ptrs;
    explicit Pool(lsp_ctx* c) : ctx(c) {}
    ~Pool() {
        for (void* p : ptrs) dev_free(ctx, p);
    }
    template <class T>
    int get(T** p, size_t bytes) {
        int rc = dev_alloc(ctx, (void**)p, bytes);
        if (rc == LSP_OK) ptrs.push_back(*p);
        return rc;
    }
};

enum { S_PUB0, S_PUB1, S_LOGN, S_ALPHA, S_ZETA, S_ZETA_NEXT, S_ALPHA_FRI, S_GEN, S_BETA, S_CHUNK_SHIFT, S_CHUNK_PT = S_CHUNK_SHIFT + 8,
       S_OPEN = S_CHUNK_PT + 8, S_COUNT = S_OPEN + 8 };

struct RankState {
    int rank = 0;
    Fr* sc = nullptr;
    DevChallenger* ch = nullptr;
    Fr* proof = nullptr;
    Fr *lde_t = nullptr, *lde_next = nullptr, *dig_t = nullptr, *top_t = nullptr;
    const Fr** cols_t = nullptr;
    Fr *chunks = nullptr, *coef_q = nullptr, *lde_q = nullptr, *dig_q = nullptr, *top_q = nullptr;
    const Fr** cols_q = nullptr;
    Fr* inv_den[2] = {nullptr, nullptr};
    Fr *folded = nullptr, *fri_dig = nullptr, *fri_top = nullptr, *tail = nullptr, *tail_dig = nullptr;
    uint32_t* idx = nullptr;
    std::vector<FriRoundS> rounds;
};

// top of a sharded tree: layer 0 = G roots already in `top`; fills the remaining G-1 digests
int merkle_top(lsp_ctx* ctx, Fr* top, int g) {
    if (g <= 1) return LSP_OK;
    // layers over the roots are plain compress layers: reuse the pair-tree builder on the G roots
    return merkle_build_pairs(ctx, top, size_t(g), top + g);
}

}  // namespace

extern "C" int lsp_nccl_unique_id(uint8_t out[128]) {
    NcclApi* a = nccl_api();
    if (!a || !out) return LSP_ERR_COMM;
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != ncclSuccess) return LSP_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(out, &id, 128);
    return LSP_OK;
}

extern "C" int lsp_comm_init_nccl(lsp_ctx* ctx, int rank, int world, const uint8_t unique_id[128], lsp_comm** out) {
    if (!ctx || !out || !unique_id || world < 1 || rank < 0 || rank >= world || (world & (world - 1))) return LSP_ERR_PARAM;
    NcclApi* a = nccl_api();
    if (!a) return set_err(ctx, LSP_ERR_COMM, "libnccl.so.2 could not be loaded");
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    lsp_comm* cm = new lsp_comm();
    cm->ctx = ctx;
    cm->world = world;
    cm->rank = rank;
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    ncclResult_t r = a->CommInitRank((ncclComm_t*)&cm->nccl, world, id, rank);
    if (r != ncclSuccess) {
        delete cm;
        return set_err(ctx, LSP_ERR_COMM, "ncclCommInitRank: %s", a->GetErrorString ? a->GetErrorString(r) : "error");
    }
    *out = cm;
    return LSP_OK;
}

extern "C" int lsp_comm_init_local(lsp_ctx* ctx, int world, lsp_comm** out) {
    if (!ctx || !out || world < 1 || (world & (world - 1))) return LSP_ERR_PARAM;
    lsp_comm* cm = new lsp_comm();
    cm->ctx = ctx;
    cm->world = world;
    cm->local = true;
    *out = cm;
    return LSP_OK;
}

extern "C" void lsp_comm_destroy(lsp_comm* cm) {
    if (!cm) return;
    if (((ncclComm_t)cm->nccl) && nccl_api()) nccl_api()->CommDestroy(((ncclComm_t)cm->nccl));
    cudaFree(cm->up_stage);
    cudaFree(cm->up_mat);
    delete cm;
}

extern "C" int lsp_prove_air_sharded_dev(lsp_comm* cm, const lsp_fri_config* fri, const lsp_mat* tr, const lsp_lookup_air_cfg* lookups,
                                         int n_lookups, const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                                         uint64_t* proof_out, size_t proof_words, float* timings_ms_out) {
    if (!cm || !fri || !tr || !publics || !proof_out || n_lookups < 0 || n_cfgs < 0 || n_lookups + n_cfgs <= 0) return LSP_ERR_PARAM;
    if ((n_cfgs && !cfgs) || (n_lookups && !lookups)) return LSP_ERR_PARAM;
    lsp_ctx* ctx = cm->ctx;
    if (!ctx->p2_set) return set_err(ctx, LSP_ERR_STATE, "lsp_set_poseidon2 has not been called");
    const size_t n = tr->rows, W = tr->width;
    if (!is_pow2(n)) return set_err(ctx, LSP_ERR_PARAM, "trace height %zu is not a power of two (prove would panic)", n);
    const int G = cm->world, log_g = ilog2(size_t(G));
    const int log_n = ilog2(n), log_q = lsp_air_log_quotient_degree_cfg(lookups, n_lookups, cfgs, n_cfgs), q = 1 << log_q;
    const int log_b = int(fri->log_blowup), log_l = log_n + log_b;
    LSP_TRY(check_fri_config(ctx, fri, log_n, log_q));
    // More ranks than cosets: a rank owns a FRACTION 2^-log_s of one coset (row block of N); its rows are the sub-coset
    // the first log_s DIF stages would hand it (coset_evaluate_subblock).  8 rows is the alignment of the 1/(x - z) tables.
    const int log_s = log_g > log_b ? log_g - log_b : 0;
    if (log_s > 0 && (log_n - log_s) < 3) return set_err(ctx, LSP_ERR_PARAM, "%d ranks are too many for 2^%d x 2^%d rows", G, log_n, log_b);
    const size_t need = lsp_proof_words(log_n, uint32_t(W), log_q, fri);
    if (proof_words < need) return set_err(ctx, LSP_ERR_PARAM, "proof buffer too small: %zu < %zu words", proof_words, need);
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));

    const size_t L = size_t(1) << log_l, Lr = L >> log_g;
    const int Bs = (1 << log_b) >> log_g;  // whole cosets (row blocks of N) per rank; 0 when a rank owns a fraction of one
    // this rank's rows of a committed matrix, from its coefficients: whole cosets, or the sub-coset of one
    auto lde_local = [&](const Fr* coeffs, size_t width, const Fr* shift_dev, int rank, Fr* out) -> int {
        if (log_s == 0) return coset_evaluate_blocks(ctx, coeffs, n, width, log_b, shift_dev, rank * Bs, Bs, out, Lr);
        return coset_evaluate_subblock(ctx, coeffs, n, width, log_b, shift_dev, rank >> log_s, log_s, rank & ((1 << log_s) - 1), out, Lr);
    };
    const int n_rounds = log_n - int(fri->log_final_poly_len);
    const int log_f = log_b + int(fri->log_final_poly_len);
    const size_t proof_elems = need / 4;
    const size_t hdr_elems = 2 + 2 * W + q + n_rounds + (size_t(1) << log_f) + 1;
    const size_t per_query = (proof_elems - hdr_elems) / fri->num_queries;
    // a commit-phase round stays sharded while every rank keeps at least this many elements
    const size_t FRI_MIN_LOCAL = 1024;

    std::vector<int> ranks;
    if (cm->local)
        for (int r = 0; r < G; r++) ranks.push_back(r);
    else
        ranks.push_back(cm->rank);
    const size_t H = ranks.size();
    Pool P(ctx);
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

    // ---- replicated inputs: trace, coefficients, AIR config ---------------------------------------
    PermCfgDev cfg_dev;
    void* cfg_blob = nullptr;
    LSP_TRY(upload_air_cfgs(ctx, lookups, n_lookups, cfgs, n_cfgs, W, &cfg_dev, &cfg_blob));
    P.ptrs.push_back(cfg_blob);
    Fr* coef_t = nullptr;
    // Every rank needs all coefficients, but not to compute them all: over NCCL each interpolates `per` of the columns and ONE
    // in-place all-gather over NVLink completes the matrix (W*N*32 bytes in total) instead of G redundant inverse NTTs.  The
    // buffer is padded to per*G columns so that every rank contributes the same count (the padding is never read).
    const size_t per = (W + size_t(G) - 1) / size_t(G);
    LSP_TRY(P.get(&coef_t, n * per * size_t(G) * 32));
    mark();  // 0
    ctx->phase = "commit_trace";
    if (!cm->local && G > 1) {
        const size_t c0 = std::min(W, size_t(cm->rank) * per), c1 = std::min(W, c0 + per);
        if (c1 > c0) LSP_TRY(interpolate_columns(ctx, tr->d + c0 * n, n, c1 - c0, coef_t + c0 * n));
        std::vector<const void*> send{coef_t + size_t(cm->rank) * per * n};
        std::vector<void*> recv{coef_t};
        LSP_TRY(coll_allgather(cm, ranks, send, recv, per * n * 32));   // in place: send == recv + rank * count
    } else {
        LSP_TRY(interpolate_columns(ctx, tr->d, n, W, coef_t));
    }

    std::vector<RankState> st(H);
    for (size_t i = 0; i < H; i++) {
        RankState& R = st[i];
        R.rank = ranks[i];
        LSP_TRY(P.get(&R.sc, S_COUNT * 32));
        LSP_CUDA(ctx, cudaMemcpyAsync(R.sc + S_PUB0, publics, 64, cudaMemcpyHostToDevice, ctx->stream));
        LSP_LAUNCH(ctx, k_set_small, 1, 1, 0, R.sc + S_LOGN, uint32_t(log_n));
        LSP_CUDA(ctx, cudaMemcpyAsync(R.sc + S_GEN, &ctx->fc->gen, 32, cudaMemcpyDeviceToDevice, ctx->stream));  // shift = GENERATOR / 1
        LSP_TRY(P.get(&R.ch, sizeof(DevChallenger)));
        LSP_TRY(challenger_init(ctx, R.ch));
        LSP_TRY(P.get(&R.proof, proof_elems * 32));
        LSP_CUDA(ctx, cudaMemsetAsync(R.proof, 0, proof_elems * 32, ctx->stream));
        LSP_TRY(P.get(&R.lde_t, Lr * W * 32));
        LSP_TRY(P.get(&R.dig_t, (2 * Lr - 1) * 32));
        LSP_TRY(P.get(&R.top_t, 2 * size_t(G) * 32));
        LSP_TRY(P.get(&R.cols_t, W * sizeof(Fr*)));
        // ---- commit to trace data: this rank's cosets only ------------------------------------------
        LSP_TRY(lde_local(coef_t, W, R.sc + S_GEN, R.rank, R.lde_t));
    }
    mark();  // 1
    auto commit_sharded = [&](auto lde_of, auto cols_of, auto dig_of, auto top_of, int width_cols, size_t proof_slot) -> int {
        std::vector<const void*> send(H);
        std::vector<void*> recv(H);
        for (size_t i = 0; i < H; i++) {
            RankState& R = st[i];
            LSP_LAUNCH(ctx, k_make_cols, unsigned((width_cols + 63) / 64), 64, 0, (const Fr*)lde_of(R), Lr, width_cols, cols_of(R));
            LSP_TRY(merkle_build(ctx, cols_of(R), width_cols, Lr, dig_of(R)));
            send[i] = dig_of(R) + (2 * Lr - 2);
            recv[i] = top_of(R);
        }
        LSP_TRY(coll_allgather(cm, ranks, send, recv, 32));
        for (size_t i = 0; i < H; i++) {
            RankState& R = st[i];
            LSP_TRY(merkle_top(ctx, top_of(R), G));
            LSP_CUDA(ctx, cudaMemcpyAsync(R.proof + proof_slot, top_of(R) + (2 * size_t(G) - 2), 32, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        return LSP_OK;
    };
    LSP_TRY(commit_sharded([](RankState& R) { return R.lde_t; }, [](RankState& R) { return R.cols_t; }, [](RankState& R) { return R.dig_t; },
                           [](RankState& R) { return R.top_t; }, int(W), 0));
    mark();  // 2

    // ---- transcript + quotient ---------------------------------------------------------------------------
    ctx->phase = "quotient";
    for (size_t i = 0; i < H; i++) {
        RankState& R = st[i];
        {   // observe(log_degree); observe(trace_commit); observe_slice(publics); alpha <- sample
            ObserveList ol;
            ol.n_seg = 3;
            ol.p[0] = R.sc + S_LOGN, ol.n[0] = 1;
            ol.p[1] = R.proof, ol.n[1] = 1;
            ol.p[2] = R.sc + S_PUB0, ol.n[2] = 2;
            LSP_TRY(challenger_observe_sample(ctx, R.ch, ol, nullptr, R.sc + S_ALPHA));
        }
        LSP_TRY(P.get(&R.chunks, size_t(q) * n * 32));
        // row block b of the quotient domain (b < q) holds chunk bitrev(b); its owner computes it (the blocks of one
        // rank are adjacent: one launch)
        if (log_s == 0) {
            const int b0 = R.rank * Bs, b1 = std::min(q, b0 + Bs);
            if (b0 < b1)
                LSP_TRY(quotient_permutation_range(ctx, R.lde_t, Lr, size_t(b0) * n, log_n, log_q, cfg_dev, R.sc + S_PUB0, R.sc + S_ALPHA,
                                                   size_t(b0) * n, size_t(b1 - b0) * n, R.chunks));
        } else {
            LSP_CUDA(ctx, cudaMemsetAsync(R.chunks, 0, size_t(q) * n * 32, ctx->stream));
            if ((R.rank >> log_s) < q) LSP_TRY(P.get(&R.lde_next, Lr * W * 32));
        }
    }
    if (log_s > 0) {
        // A fraction of a coset: the next row of a row is NOT among this rank's rows (adjacent trace rows are far apart in
        // bit-reversed order).  A rank holds the trace rows k = k0 (mod S) of its coset, so ALL its next rows are the rows of
        // ONE neighbour, the rank with residue k0 + 1 (same local row; rotated by one evaluation point when k0 + 1 wraps to 0):
        // the ranks of a quotient-domain coset pass their local LDE one step round the ring (NVLink send/recv: Lr*W*32 bytes,
        // a quarter of the time of evaluating p(w_N x) locally), and every rank's scattered share of a chunk is summed below.
        const int S = 1 << log_s;
        auto rank_of = [&](int block, int k0) { return block * S + int(bitrev_host(uint32_t(k0), log_s)); };
        std::vector<int> send_to(H, -1), recv_from(H, -1);
        std::vector<const void*> send(cm->local ? size_t(G) : H, nullptr);
        std::vector<void*> recv(H, nullptr);
        for (size_t i = 0; i < H; i++) {
            RankState& R = st[i];
            const int block = R.rank >> log_s, k0 = int(bitrev_host(uint32_t(R.rank & (S - 1)), log_s));
            send[cm->local ? size_t(R.rank) : i] = R.lde_t;
            if (block >= q) continue;
            send_to[i] = rank_of(block, (k0 + S - 1) & (S - 1));
            recv_from[i] = rank_of(block, (k0 + 1) & (S - 1));
            recv[i] = R.lde_next;
        }
        LSP_TRY(coll_sendrecv(cm, ranks, send_to, recv_from, send, recv, Lr * W * 32));
        for (size_t i = 0; i < H; i++) {
            RankState& R = st[i];
            const int block = R.rank >> log_s, k0 = int(bitrev_host(uint32_t(R.rank & (S - 1)), log_s));
            if (block >= q) continue;
            LSP_TRY(quotient_permutation_range(ctx, R.lde_t, Lr, size_t(R.rank) * Lr, log_n, log_q, cfg_dev, R.sc + S_PUB0, R.sc + S_ALPHA,
                                               size_t(R.rank) * Lr, Lr, R.chunks, R.lde_next, /*next_rotated=*/k0 == S - 1));
        }
    }
    if (log_s > 0) {  // disjoint shares + zeros: an integer sum assembles every chunk on every rank
        std::vector<void*> buf(H);
        for (size_t i = 0; i < H; i++) buf[i] = st[i].chunks;
        if (cm->local) {
            Fr* summed = nullptr;
            LSP_TRY(P.get(&summed, size_t(q) * n * 32));
            LSP_TRY(coll_allreduce_sum(cm, ranks, buf, size_t(q) * n * 32, summed));
            for (size_t i = 0; i < H; i++) st[i].chunks = summed;
        } else {
            LSP_TRY(coll_allreduce_sum(cm, ranks, buf, size_t(q) * n * 32, nullptr));
        }
    }
    for (int b = 0; b < q && G > 1 && log_s == 0; b++) {  // the one bulk exchange: N field elements per chunk
        const int c = int(bitrev_host(uint32_t(b), log_q));
        const int owner = b / Bs;
        std::vector<void*> buf(H);
        for (size_t i = 0; i < H; i++) buf[i] = st[i].chunks + size_t(c) * n;
        // (local mode: buf[] is indexed by hosted-rank position == global rank)
        LSP_TRY(coll_broadcast(cm, ranks, owner, buf, n * 32));
    }
    mark();  // 3

    // ---- commit to quotient poly chunks ----------------------------------------------------------------
    ctx->phase = "commit_quotient";
    for (size_t i = 0; i < H; i++) {
        RankState& R = st[i];
        LSP_TRY(P.get(&R.coef_q, size_t(q) * n * 32));
        LSP_TRY(P.get(&R.lde_q, size_t(q) * Lr * 32));
        LSP_TRY(P.get(&R.dig_q, (2 * Lr - 1) * 32));
        LSP_TRY(P.get(&R.top_q, 2 * size_t(G) * 32));
        LSP_TRY(P.get(&R.cols_q, q * sizeof(Fr*)));
        LSP_LAUNCH(ctx, k_chunk_consts, 1, 32, 0, (const FieldConsts*)ctx->fc, (const Fr*)nullptr, log_n, log_q, R.sc + S_CHUNK_SHIFT, (Fr*)nullptr, (Fr*)nullptr);
        LSP_TRY(interpolate_columns(ctx, R.chunks, n, q, R.coef_q));
        for (int c = 0; c < q; c++)
            LSP_TRY(lde_local(R.coef_q + size_t(c) * n, 1, R.sc + S_CHUNK_SHIFT + c, R.rank, R.lde_q + size_t(c) * Lr));
    }
    LSP_TRY(commit_sharded([](RankState& R) { return R.lde_q; }, [](RankState& R) { return R.cols_q; }, [](RankState& R) { return R.dig_q; },
                           [](RankState& R) { return R.top_q; }, q, 1));
    mark();  // 4

    // ---- open: opened values (replicated), reduced openings (local rows) ---------------------------
    ctx->phase = "open";
    for (size_t i = 0; i < H; i++) {
        RankState& R = st[i];
        Fr* p_local = R.proof + 2;
        Fr* p_next = p_local + W;
        Fr* p_chunks = p_next + W;
        LSP_TRY(challenger_observe_sample(ctx, R.ch, R.proof + 1, 1, nullptr, R.sc + S_ZETA));   // observe(quotient_commit); zeta <- sample
        LSP_LAUNCH(ctx, k_chunk_consts, 1, 32, 0, (const FieldConsts*)ctx->fc, (const Fr*)(R.sc + S_ZETA), log_n, log_q, (Fr*)nullptr, R.sc + S_ZETA_NEXT, R.sc + S_CHUNK_PT);
        // `TwoAdicFriPcs::open`: the pinned fork samples the batching challenge first and never observes the opened
        // values; later upstream observes them (trace at zeta, trace at zeta', each chunk at zeta) and samples after.
        if (ctx->alpha_before_openings) LSP_TRY(challenger_sample(ctx, R.ch, R.sc + S_ALPHA_FRI));
        LSP_TRY(eval_columns_at(ctx, coef_t, n, W, R.sc + S_ZETA, p_local));
        LSP_TRY(eval_columns_at(ctx, coef_t, n, W, R.sc + S_ZETA_NEXT, p_next));
        for (int c = 0; c < q; c++) LSP_TRY(eval_columns_at(ctx, R.coef_q + size_t(c) * n, n, 1, R.sc + S_CHUNK_PT + c, p_chunks + c));
        if (ctx->observe_opened_values) LSP_TRY(challenger_observe_dev(ctx, R.ch, p_local, int(2 * W + q)));
        if (!ctx->alpha_before_openings) LSP_TRY(challenger_sample(ctx, R.ch, R.sc + S_ALPHA_FRI));
        LSP_LAUNCH(ctx, k_open_scalars, 1, 1, 0, R.sc + S_ALPHA_FRI, p_local, p_next, p_chunks, int(W), q, R.sc + S_OPEN);
        LSP_TRY(P.get(&R.inv_den[0], Lr * 32));
        LSP_TRY(P.get(&R.inv_den[1], Lr * 32));
        LSP_TRY(inverse_denominators_range(ctx, R.sc + S_ZETA, 2, log_l, size_t(R.rank) * Lr, Lr, R.inv_den));
        LSP_TRY(P.get(&R.folded, 2 * Lr * 32));
        ReduceArgsS A;
        A.trace_lde = R.lde_t;
        A.rows = Lr;
        A.width = int(W);
        A.quot_lde = R.lde_q;
        A.q = q;
        A.alpha = R.sc + S_ALPHA_FRI;
        A.s = R.sc + S_OPEN;
        A.e_zeta = R.inv_den[0];
        A.e_next = R.inv_den[1];
        A.out = R.folded;
        LSP_LAUNCH(ctx, k_reduce_openings_s, grid_for(ctx, Lr, 128), 128, 0, A);
    }
    mark();  // 5

    // ---- FRI commit phase: sharded rounds, then one all-gather and replicated rounds ---------------
    ctx->phase = "fri_commit";
    for (size_t i = 0; i < H; i++) {
        LSP_TRY(P.get(&st[i].fri_dig, 2 * Lr * 32));
        LSP_TRY(P.get(&st[i].fri_top, size_t(n_rounds + 1) * 2 * G * 32));
        st[i].rounds.resize(n_rounds);
    }
    size_t len = L;
    int r = 0;
    {
        std::vector<Fr*> cur(H), dig(H);
        for (size_t i = 0; i < H; i++) {
            cur[i] = st[i].folded;
            dig[i] = st[i].fri_dig;
        }
        for (; r < n_rounds && G > 1 && (len >> log_g) >= 2 * FRI_MIN_LOCAL; r++) {
            const size_t local = len >> log_g;
            std::vector<const void*> send(H);
            std::vector<void*> recv(H);
            for (size_t i = 0; i < H; i++) {
                LSP_TRY(merkle_build_pairs(ctx, cur[i], local, dig[i]));
                send[i] = dig[i] + (local - 2);
                recv[i] = st[i].fri_top + size_t(r) * 2 * G;
            }
            LSP_TRY(coll_allgather(cm, ranks, send, recv, 32));
            for (size_t i = 0; i < H; i++) {
                RankState& R = st[i];
                Fr* top = R.fri_top + size_t(r) * 2 * G;
                LSP_TRY(merkle_top(ctx, top, G));
                const Fr* root = top + (2 * size_t(G) - 2);
                LSP_TRY(challenger_observe_sample(ctx, R.ch, root, 1, R.proof + 2 + 2 * W + q + r, R.sc + S_BETA));
                Fr* nxt = cur[i] + local;
                LSP_TRY(fri_fold_range(ctx, cur[i], len, size_t(R.rank) * (local / 2), local / 2, R.sc + S_BETA, nxt));
                R.rounds[r].vec = cur[i];
                R.rounds[r].tree.local = dig[i];
                R.rounds[r].tree.top = top;
                R.rounds[r].tree.log_h = uint32_t(ilog2(len) - 1);
                R.rounds[r].sharded = 1;
                dig[i] += local - 1;
                cur[i] = nxt;
            }
            len >>= 1;
        }
        // gather what is left and finish identically on every rank
        const size_t local = len >> log_g;
        std::vector<const void*> send(H);
        std::vector<void*> recv(H);
        if (G > 1) {
            for (size_t i = 0; i < H; i++) {
                LSP_TRY(P.get(&st[i].tail, 2 * len * 32));
                LSP_TRY(P.get(&st[i].tail_dig, 2 * len * 32));
                send[i] = cur[i];
                recv[i] = st[i].tail;
            }
            LSP_TRY(coll_allgather(cm, ranks, send, recv, local * 32));
        } else {  // one rank: nothing to gather, the rounds continue in place
            st[0].tail = cur[0];
            st[0].tail_dig = dig[0];
        }
        for (size_t i = 0; i < H; i++) {
            RankState& R = st[i];
            Fr* c = R.tail;
            Fr* d = R.tail_dig;
            size_t ln = len;
            for (int rr = r; rr < n_rounds; rr++) {
                LSP_TRY(merkle_build_pairs(ctx, c, ln, d));
                const Fr* root = d + (ln - 2);
                LSP_TRY(challenger_observe_sample(ctx, R.ch, root, 1, R.proof + 2 + 2 * W + q + rr, R.sc + S_BETA));
                LSP_TRY(fri_fold(ctx, c, ln, R.sc + S_BETA, c + ln));
                R.rounds[rr].vec = c;
                R.rounds[rr].tree.local = d;
                R.rounds[rr].tree.top = nullptr;
                R.rounds[rr].tree.log_h = uint32_t(ilog2(ln) - 1);
                R.rounds[rr].sharded = 0;
                d += ln - 1;
                c += ln;
                ln >>= 1;
            }
            Fr* p_final = R.proof + 2 + 2 * W + q + n_rounds;
            LSP_LAUNCH(ctx, k_final_poly, 1, unsigned(ln < 32 ? 32 : ln), 0, (const FieldConsts*)ctx->fc, (const Fr*)c, log_f, host_pow2_inverse(log_f), p_final);
            LSP_TRY(challenger_observe_dev(ctx, R.ch, p_final, int(ln)));
        }
    }
    mark();  // 6

    // ---- grind + queries ---------------------------------------------------------------------------------
    ctx->phase = "fri_query";
    std::vector<void*> qbuf(H);
    for (size_t i = 0; i < H; i++) {
        RankState& R = st[i];
        Fr* p_pow = R.proof + hdr_elems - 1;
        LSP_TRY(challenger_grind(ctx, R.ch, int(fri->proof_of_work_bits), p_pow));
        LSP_TRY(P.get(&R.idx, fri->num_queries * 4));
        LSP_TRY(challenger_sample_bits(ctx, R.ch, log_l, int(fri->num_queries), R.idx));
        QueryArgsS A;
        A.idx = R.idx;
        A.rank = R.rank;
        A.log_g = log_g;
        A.trace_lde = R.lde_t;
        A.rows = Lr;
        A.width = int(W);
        A.trace_tree = ShardTree{R.dig_t, R.top_t, uint32_t(log_l)};
        A.quot_lde = R.lde_q;
        A.q = q;
        A.quot_tree = ShardTree{R.dig_q, R.top_q, uint32_t(log_l)};
        A.log_l = log_l;
        for (int rr = 0; rr < n_rounds; rr++) A.rounds[rr] = R.rounds[rr];
        A.n_rounds = n_rounds;
        A.out = R.proof + hdr_elems;
        A.per_query = per_query;
        LSP_LAUNCH(ctx, k_query_gather_s, fri->num_queries, 128, 0, A);
        qbuf[i] = R.proof + hdr_elems;
    }
    const size_t q_bytes = (proof_elems - hdr_elems) * 32;
    Fr* out_dev = st[0].proof;
    if (G > 1) {
        if (cm->local) {
            Fr* summed = nullptr;
            LSP_TRY(P.get(&summed, q_bytes));
            LSP_TRY(coll_allreduce_sum(cm, ranks, qbuf, q_bytes, summed));
            LSP_CUDA(ctx, cudaMemcpyAsync(st[0].proof + hdr_elems, summed, q_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
            LSP_TRY(coll_allreduce_sum(cm, ranks, qbuf, q_bytes, nullptr));
        }
    }
    mark();  // 7
    LSP_CUDA(ctx, cudaMemcpyAsync(proof_out, out_dev, proof_elems * 32, cudaMemcpyDeviceToHost, ctx->stream));
    mark();  // 8
    ctx->phase = "";
    LSP_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, &st[0].ch->overflow, 4, cudaMemcpyDeviceToHost, ctx->stream));
    LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*(volatile int*)ctx->pinned) return set_err(ctx, LSP_ERR_STATE, "challenger input buffer overflow");
    if (timings_ms_out)
        for (int i = 0; i < 8; i++) cudaEventElapsedTime(&timings_ms_out[i], ev[i], ev[i + 1]);
    return LSP_OK;
}

// Host-trace entry point.  Over NCCL every rank uploads only ITS 1/G of the rows -- a contiguous slice of the
// row-major host matrix, written at its place in a full-size row-major device buffer -- and an in-place
// all-gather over NVLink completes the matrix: the trace crosses PCIe once in total instead of once per rank
// (8 x 1 GiB at 2^22 rows).  Staging buffer and matrix belong to the communicator and are reused from call to
// call, so this path leaves the stream-ordered pool -- and with it the prove's allocation pattern -- untouched.
extern "C" int lsp_prove_air_sharded(lsp_comm* cm, const lsp_fri_config* fri, const uint64_t* trace, size_t rows, size_t width,
                                     const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* cfgs, int n_cfgs,
                                     const uint64_t publics[2][4], uint64_t* proof_out, size_t proof_words, float* timings_ms_out) {
    if (!cm || !trace) return LSP_ERR_PARAM;
    lsp_ctx* ctx = cm->ctx;
    const size_t G = size_t(cm->world);
    if (cm->local || G == 1 || rows % G != 0 || rows / G < 64) {
        lsp_mat* m = nullptr;
        LSP_TRY(lsp_mat_upload(ctx, trace, rows, width, &m));
        int rc = lsp_prove_air_sharded_dev(cm, fri, m, lookups, n_lookups, cfgs, n_cfgs, publics, proof_out, proof_words, timings_ms_out);
        lsp_mat_free(ctx, m);
        return rc;
    }
    LSP_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t elems = rows * width, slice = rows / G;
    if (cm->up_elems < elems) {
        LSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(cm->up_stage);
        cudaFree(cm->up_mat);
        cm->up_stage = cm->up_mat = nullptr;
        cm->up_elems = 0;
        LSP_CUDA(ctx, cudaMalloc((void**)&cm->up_stage, elems * 32));
        LSP_CUDA(ctx, cudaMalloc((void**)&cm->up_mat, elems * 32));
        cm->up_elems = elems;
    }
    Fr* mine = cm->up_stage + size_t(cm->rank) * slice * width;
    LSP_CUDA(ctx, cudaMemcpyAsync(mine, trace + size_t(cm->rank) * slice * width * 4, slice * width * 32, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int> ranks{cm->rank};
    std::vector<const void*> send{mine};
    std::vector<void*> recv{cm->up_stage};
    LSP_TRY(coll_allgather(cm, ranks, send, recv, slice * width * 32));   // in place: send == recv + rank * count
    LSP_TRY(rowmajor_to_colmajor(ctx, cm->up_stage, rows, width, cm->up_mat));
    lsp_mat view;
    view.d = cm->up_mat;
    view.rows = rows;
    view.width = width;
    view.owns = false;
    return lsp_prove_air_sharded_dev(cm, fri, &view, lookups, n_lookups, cfgs, n_cfgs, publics, proof_out, proof_words, timings_ms_out);
}

extern "C" int lsp_prove_permutation_sharded_dev(lsp_comm* cm, const lsp_fri_config* fri, const lsp_mat* tr, const lsp_perm_air_cfg* cfgs,
                                                 int n_cfgs, const uint64_t publics[2][4], uint64_t* proof_out, size_t proof_words,
                                                 float* timings_ms_out) {
    if (!cfgs || n_cfgs <= 0) return LSP_ERR_PARAM;
    return lsp_prove_air_sharded_dev(cm, fri, tr, nullptr, 0, cfgs, n_cfgs, publics, proof_out, proof_words, timings_ms_out);
}

extern "C" int lsp_prove_permutation_sharded(lsp_comm* cm, const lsp_fri_config* fri, const uint64_t* trace, size_t rows, size_t width,
                                             const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4], uint64_t* proof_out,
                                             size_t proof_words, float* timings_ms_out) {
    if (!cfgs || n_cfgs <= 0) return LSP_ERR_PARAM;
    return lsp_prove_air_sharded(cm, fri, trace, rows, width, nullptr, 0, cfgs, n_cfgs, publics, proof_out, proof_words, timings_ms_out);
}
