// Width-3 Poseidon2 over BLS12-377 Fr, state held in registers.
//
// Device-side replacement for the reference's
//   Perm     = Poseidon2Bls12337<3>                      (bin/src/config.rs:11, bin/src/main.rs:49)
//   Hash     = PaddingFreeSponge<Perm,3,2,1>             (bin/src/config.rs:12)
//   Compress = CompressionFunctionFromHasher<Hash,2,1>   (bin/src/config.rs:17)
// Round constants, round counts, S-box degree and the internal diagonal are
// runtime parameters (lsp_set_poseidon2): the reference draws the constants
// at start-up (bin/src/main.rs:49) and the fork-only crate that fixes the rest
// is not available, so nothing is assumed.
//
// Two shapes of the same permutation (bit-identical results):
//   p2_permute      one thread per permutation, state in a per-thread shared-memory slot.
//                   Throughput shape: the 46 S-boxes run through ONE inlined S-box body inside a
//                   flat loop, so the kernel is ~12 KiB of code (fits the 32 KiB L1.5 I-cache),
//                   makes no calls and passes nothing through local memory.
//   p2_permute_tri  three adjacent lanes per permutation, one state word each.  Latency shape
//                   for the small Merkle layers and the Fiat-Shamir transcript, whose cost is a
//                   chain of dependent permutations: a full round's three S-boxes run side by
//                   side, so a permutation is 30 S-box latencies instead of 46.
//
// Values are kept unreduced ("lazy", a small multiple of r) between the linear layers and the
// S-boxes; the bounds are stated at each step.  2^256 / r = 13.7, so anything below 13 r fits.
#pragma once
#include "fr.cuh"

namespace lsp {

constexpr int P2_WIDTH = 3;
constexpr int P2_MAX_HALF_F = 8;
constexpr int P2_MAX_P = 64;

// Passed by value as a __grid_constant__ kernel parameter (constant bank).
struct P2Params {
    int half_f;    // rounds_f / 2
    int rounds_p;
    int sbox_d;
    int diag_kind;  // 0: generic diag, 1: (1,1,2)
    Fr ext_initial[P2_MAX_HALF_F][3];
    Fr ext_terminal[P2_MAX_HALF_F][3];
    Fr internal[P2_MAX_P];
    Fr diag_m1[3];
};

// ---- lazy helpers ---------------------------------------------------------------------------
__device__ __forceinline__ Fr fr_add_lazy(const Fr& a, const Fr& b) {  // no reduction; caller bounds the sum below 2^256
    Fr r;
    u256_add(r.l, a.l, b.l);
    return r;
}
// a -> a - K*r when a >= K*r  (K in {1, 2, 4})
template <int K>
__device__ __forceinline__ void fr_cond_sub(Fr& a) {
    const uint32_t k1[8] = {LSP_P0, LSP_P1, LSP_P2, LSP_P3, LSP_P4, LSP_P5, LSP_P6, LSP_P7};
    const uint32_t k2[8] = {0x00000002u, 0x14230000u, 0xa0000002u, 0xb354edfdu, 0xb86f6002u, 0xc1689a3cu, 0x34594aacu, 0x2556cabdu};
    const uint32_t k4[8] = {0x00000004u, 0x28460000u, 0x40000004u, 0x66a9dbfbu, 0x70dec005u, 0x82d13479u, 0x68b29559u, 0x4aad957au};
    const uint32_t* k = K == 1 ? k1 : (K == 2 ? k2 : k4);
    uint32_t t[8];
    uint32_t bw = u256_sub(t, a.l, k);
#pragma unroll
    for (int i = 0; i < 8; i++) a.l[i] = bw ? a.l[i] : t[i];
}
// a < 4r -> canonical
__device__ __forceinline__ void fr_canon4(Fr& a) {
    fr_cond_sub<2>(a);
    fr_cond_sub<1>(a);
}
// a < 8r -> canonical
__device__ __forceinline__ void fr_canon8(Fr& a) {
    fr_cond_sub<4>(a);
    fr_cond_sub<2>(a);
    fr_cond_sub<1>(a);
}

// x^D for x < 5r (x < 4r when D == 3); canonical result.  With out < a*b/2^256 + r for the lazy
// product: x2 < 2.83r, x4 < 1.59r, x8 < 1.19r, x16 < 1.11r and every final product is < 1.6r, so
// one conditional subtraction finishes.
template <int D>
__device__ __forceinline__ Fr p2_sbox(const Fr& x) {
    Fr y;
    if (D == 3) {
        y = fr_mul_lazy(fr_sqr_lazy(x), x);
    } else if (D == 5) {
        Fr x2 = fr_sqr_lazy(x);
        y = fr_mul_lazy(fr_sqr_lazy(x2), x);
    } else if (D == 7) {
        Fr x2 = fr_sqr_lazy(x);
        Fr x4 = fr_sqr_lazy(x2);
        y = fr_mul_lazy(fr_mul_lazy(x4, x2), x);
    } else if (D == 11) {
        Fr x2 = fr_sqr_lazy(x);
        Fr x8 = fr_sqr_lazy(fr_sqr_lazy(x2));
        y = fr_mul_lazy(fr_mul_lazy(x8, x2), x);
    } else {  // 17
        Fr x16 = fr_sqr_lazy(fr_sqr_lazy(fr_sqr_lazy(fr_sqr_lazy(x))));
        y = fr_mul_lazy(x16, x);
    }
    fr_reduce_once(y);
    return y;
}
// S-box input: s (< 4r) + round constant (< r).
template <int D>
__device__ __forceinline__ Fr p2_sbox_in(const Fr& s, const Fr& c) {
    Fr x = fr_add_lazy(s, c);  // < 5r
    if (D == 3) fr_cond_sub<4>(x);  // x^3 needs x < 4r for its one-subtraction finish
    return x;
}

// M_E = circ(2,1,1): s_i += s0+s1+s2.  Canonical in -> < 4r out.
__device__ __forceinline__ void p2_ext_linear_lazy(Fr& s0, Fr& s1, Fr& s2) {
    Fr t = fr_add_lazy(fr_add_lazy(s0, s1), s2);
    s0 = fr_add_lazy(s0, t);
    s1 = fr_add_lazy(s1, t);
    s2 = fr_add_lazy(s2, t);
}
// Internal layer s_i <- diag_i s_i + (s0+s1+s2), canonical in; s0 leaves lazy (< 4r, it goes straight
// into the next S-box), s1 and s2 canonical (they wait out the partial rounds).
__device__ __forceinline__ void p2_int_linear(const P2Params& P, Fr& s0, Fr& s1, Fr& s2) {
    Fr t = fr_add_lazy(fr_add_lazy(s0, s1), s2);  // < 3r
    if (P.diag_kind == 1) {                        // diag - 1 = (1, 1, 2): add-only
        s0 = fr_add_lazy(s0, t);
        s1 = fr_add_lazy(s1, t);                   // < 4r
        fr_canon4(s1);
        s2 = fr_add_lazy(fr_add_lazy(s2, s2), t);  // < 5r
        fr_canon8(s2);
    } else {
        fr_canon4(t);
        s0 = fr_add(fr_mul(s0, P.diag_m1[0]), t);
        s1 = fr_add(fr_mul(s1, P.diag_m1[1]), t);
        s2 = fr_add(fr_mul(s2, P.diag_m1[2]), t);
    }
}

// ---- one thread per permutation -------------------------------------------------------------
// Canonical in, canonical out.  Flat loop over the 3*RF + RP S-box applications with a single
// inlined S-box body.  The three state words live in a per-thread slot of SHARED memory
// (6 x uint4, stride NT = threads per block => conflict-free 128-bit accesses): the word an
// S-box acts on is then a dynamic address, not a register choice.  With the state in registers
// the same loop needs ~40 selects/moves per S-box, which ptxas emits as IMAD.MOV on the
// FMA-heavy pipe -- the pipe this kernel saturates -- and 24 more live registers.
template <int NT>
struct P2Slot {
    uint4* base;  // &slot_array[threadIdx.x]; slot_array holds 6 * NT uint4
    __device__ __forceinline__ Fr load(int w) const {
        const uint4 a = base[(2 * w) * NT], b = base[(2 * w + 1) * NT];
        Fr r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void store(int w, const Fr& v) const {
        base[(2 * w) * NT] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        base[(2 * w + 1) * NT] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
};
#define LSP_P2_SLOT_DECL(NT) __shared__ uint4 lsp_p2_slots[6 * (NT)]
#define LSP_P2_SLOT(NT) (lsp::P2Slot<NT>{lsp_p2_slots + threadIdx.x})

template <int D, int NT>
__device__ __forceinline__ void p2_permute(const P2Params& P, Fr& s0, Fr& s1, Fr& s2, const P2Slot<NT> st) {
    p2_ext_linear_lazy(s0, s1, s2);
    st.store(0, s0);
    st.store(1, s1);
    st.store(2, s2);
    const int nf = 3 * P.half_f, np = P.rounds_p, total = 2 * nf + np;
    const Fr* ci = &P.ext_initial[0][0];
    const Fr* ct = &P.ext_terminal[0][0];
    int sel = 0;  // word of a full round (0, 1, 2); partial rounds always act on word 0
#pragma unroll 1
    for (int t = 0; t < total; t++) {
        const bool partial = t >= nf && t < nf + np;
        const Fr* cp = t < nf ? ci + t : (partial ? &P.internal[t - nf] : ct + (t - nf - np));
        const Fr y = p2_sbox<D>(p2_sbox_in<D>(st.load(sel), *cp));
        if (partial) {
            Fr a = y, b = st.load(1), c = st.load(2);
            if (t == nf) {  // words 1, 2 arrive lazy from the last full round
                fr_canon4(b);
                fr_canon4(c);
            }
            p2_int_linear(P, a, b, c);
            st.store(0, a);
            st.store(1, b);
            st.store(2, c);
        } else if (sel == 2) {
            Fr a = st.load(0), b = st.load(1), c = y;
            p2_ext_linear_lazy(a, b, c);
            st.store(0, a);
            st.store(1, b);
            st.store(2, c);
            sel = 0;
        } else {
            st.store(sel, y);
            sel++;
        }
    }
    s0 = st.load(0);
    s1 = st.load(1);
    s2 = st.load(2);
    fr_canon4(s0);
    fr_canon4(s1);
    fr_canon4(s2);
}

// compress(l, r) = perm([l, r, 0])[0]
template <int D, int NT>
__device__ __forceinline__ Fr p2_compress(const P2Params& P, const Fr& l, const Fr& r, const P2Slot<NT> st) {
    Fr s0 = l, s1 = r, s2 = fr_zero();
    p2_permute<D, NT>(P, s0, s1, s2, st);
    return s0;
}

// ---- three lanes per permutation ------------------------------------------------------------
// Lane 3k+w of a warp holds word w of permutation k (10 permutations per warp; lanes 30 and 31
// idle but must execute the call: the shuffles name the full warp).  `w` = lane % 3.
__device__ __forceinline__ Fr p2_tri_sum(const Fr& s, int base) {
    Fr a, b, c;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a.l[i] = __shfl_sync(0xffffffffu, s.l[i], base);
        b.l[i] = __shfl_sync(0xffffffffu, s.l[i], base + 1);
        c.l[i] = __shfl_sync(0xffffffffu, s.l[i], base + 2);
    }
    return fr_add_lazy(fr_add_lazy(a, b), c);
}
template <int D>
__device__ __forceinline__ void p2_permute_tri(const P2Params& P, Fr& s, int w, int base) {
    s = fr_add_lazy(s, p2_tri_sum(s, base));  // < 4r
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
        for (int r = 0; r < P.half_f; r++) {
            const Fr c = phase == 0 ? P.ext_initial[r][w] : P.ext_terminal[r][w];
            s = p2_sbox<D>(p2_sbox_in<D>(s, c));
            s = fr_add_lazy(s, p2_tri_sum(s, base));
        }
        if (phase == 0) {
#pragma unroll 1
            for (int r = 0; r < P.rounds_p; r++) {
                // only word 0 takes the S-box; the other lanes run it on their own word and drop the result
                const Fr y = p2_sbox<D>(p2_sbox_in<D>(s, P.internal[r]));
                if (w == 0) {
                    s = y;
                } else if (r == 0) {
                    fr_canon4(s);  // words 1, 2 arrive lazy from the last full round
                }
                const Fr t = p2_tri_sum(s, base);  // < 3r
                if (P.diag_kind == 1) {
                    Fr u = w == 2 ? fr_add_lazy(s, s) : s;
                    s = fr_add_lazy(u, t);  // word 0: < 4r, stays lazy; words 1, 2: < 5r
                    if (w != 0) fr_canon8(s);
                } else {
                    Fr tt = t;
                    fr_canon4(tt);
                    s = fr_add(fr_mul(s, P.diag_m1[w]), tt);
                }
            }
        }
    }
    fr_canon4(s);
}

// Dispatch a templated launch on the runtime S-box degree.
#define LSP_DISPATCH_SBOX(d, ...)                  \
    switch (d) {                                   \
        case 3: { constexpr int D = 3; __VA_ARGS__; break; }   \
        case 5: { constexpr int D = 5; __VA_ARGS__; break; }   \
        case 7: { constexpr int D = 7; __VA_ARGS__; break; }   \
        case 11: { constexpr int D = 11; __VA_ARGS__; break; } \
        case 17: { constexpr int D = 17; __VA_ARGS__; break; } \
        default: return LSP_ERR_PARAM;             \
    }

}  // namespace lsp
