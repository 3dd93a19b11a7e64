/* lsp_oracle.c -- CPU restatement of the reference's prover loop.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the C twin of oracle/ Python modules: a multi-threaded (OpenMP) port of the
 * algorithm distributed-lab/linea-stark-prover runs through
 * `p3_uni_stark::prove` / `verify` (reference bin/src/main.rs:80-96) with the
 * types of bin/src/config.rs:9-25, on the permutation AIR of
 * air/src/lib.rs:116-167 and the witness of trace/src/permutation.rs:24-93.
 * It exists to (1) check the CUDA path bit-for-bit at the full 2^19-row size,
 * where the Python oracle is too slow, and (2) be timed as the "port" CPU
 * baseline.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product never links or dlopens it.
 *
 * PARITY UNPINNED: the reference's arithmetic lives in an un-vendored git
 * dependency (Plonky3 fork rev f888f90, Cargo.lock:505) and no Rust toolchain
 * exists here; the algorithms follow the published Plonky3 of that era
 * (SURVEY.md Appendix A).  Field constants are pinned to arkworks' published
 * values through the Python oracle's tests; this file is pinned to the Python
 * oracle by tests/test_oracle_c.py (whole-proof equality).
 *
 * Where the reference uses a specific CPU algorithm (barycentric
 * `interpolate_coset`, batch inversion for selectors and inverse denominators,
 * per-row inversion in witness generation) the same algorithm is used here, so
 * the timing is representative of the reference's CPU work.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fr;

static const uint64_t P[4] = {0x0a11800000000001ull, 0x59aa76fed0000001ull, 0x60b44d1e5c37b001ull, 0x12ab655e9a2ca556ull};
static const uint64_t NINV = 0x0a117fffffffffffull;                       /* -r^-1 mod 2^64 */
static const fr ONE = {{0x7d1c7ffffffffff3ull, 0x7257f50f6ffffff2ull, 0x16d81575512c0feeull, 0x0d4bda322bbb9a9dull}};
static const fr R2 = {{0x25d577bab861857bull, 0xcc2c27b58860591full, 0xa7cc008fe5dc8593ull, 0x011fdae7eff1c939ull}};
static const fr ZERO = {{0, 0, 0, 0}};
#define TWO_ADICITY 47

/* ---- field ------------------------------------------------------------- */
static inline int geq_p(const uint64_t* t) {
    for (int i = 3; i >= 0; i--) { if (t[i] > P[i]) return 1; if (t[i] < P[i]) return 0; }
    return 1;
}
static inline void sub_p(uint64_t* t) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - P[i] - b; t[i] = (uint64_t)d; b = (d >> 64) & 1; }
}
static inline fr fr_add(fr a, fr b) {
    fr r; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (geq_p(r.l)) sub_p(r.l);
    return r;
}
static inline fr fr_sub(fr a, fr b) {
    fr r; u128 bw = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a.l[i] - b.l[i] - bw; r.l[i] = (uint64_t)d; bw = (d >> 64) & 1; }
    if (bw) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.l[i] + P[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
    return r;
}
static inline fr fr_mul(fr a, fr b) {
    uint64_t t[4] = {0, 0, 0, 0}, t4 = 0;
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t4; uint64_t hi = (uint64_t)c; uint64_t hi2 = (uint64_t)(c >> 64);
        uint64_t m = t[0] * NINV;
        c = (u128)m * P[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += hi; t[3] = (uint64_t)c; t4 = (uint64_t)(c >> 64) + hi2;
    }
    fr r; memcpy(r.l, t, 32);
    if (t4 || geq_p(r.l)) sub_p(r.l);
    return r;
}
static inline fr fr_sqr(fr a) { return fr_mul(a, a); }
static inline int fr_is_zero(fr a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
static inline int fr_eq(fr a, fr b) { return a.l[0] == b.l[0] && a.l[1] == b.l[1] && a.l[2] == b.l[2] && a.l[3] == b.l[3]; }
static inline fr fr_neg(fr a) { return fr_sub(ZERO, a); }
static fr fr_pow_u64(fr a, uint64_t e) {
    fr r = ONE;
    while (e) { if (e & 1) r = fr_mul(r, a); a = fr_sqr(a); e >>= 1; }
    return r;
}
static fr fr_inv(fr a) { /* a^(r-2) */
    uint64_t e[4] = {P[0] - 2, P[1], P[2], P[3]};
    fr r = ONE;
    for (int i = 3; i >= 0; i--) for (int b = 63; b >= 0; b--) { r = fr_sqr(r); if ((e[i] >> b) & 1) r = fr_mul(r, a); }
    return r;
}
static inline fr fr_from_u64(uint64_t v) { fr x = {{v, 0, 0, 0}}; return fr_mul(x, R2); }
static inline fr fr_canonical(fr a) { fr one = {{1, 0, 0, 0}}; return fr_mul(a, one); }
static inline fr fr_halve(fr a) {
    uint64_t t[4]; memcpy(t, a.l, 32); uint64_t top = 0;
    if (t[0] & 1) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)t[i] + P[i]; t[i] = (uint64_t)c; c >>= 64; } top = (uint64_t)c; }
    fr r; for (int i = 0; i < 3; i++) r.l[i] = (t[i] >> 1) | (t[i + 1] << 63);
    r.l[3] = (t[3] >> 1) | (top << 63);
    return r;
}
static fr GEN, GEN_INV, ROOT47, HALF;
static int g_init = 0;
static void init_consts(void) {
    if (g_init) return;
    GEN = fr_from_u64(22);
    GEN_INV = fr_inv(GEN);
    /* (r-1)/2^47 */
    uint64_t e[4] = {P[0] - 1, P[1], P[2], P[3]};
    for (int s = 0; s < TWO_ADICITY; s++) { for (int i = 0; i < 3; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 63); e[3] >>= 1; }
    fr r = ONE;
    for (int i = 3; i >= 0; i--) for (int b = 63; b >= 0; b--) { r = fr_sqr(r); if ((e[i] >> b) & 1) r = fr_mul(r, GEN); }
    ROOT47 = r;
    HALF = fr_halve(ONE);
    g_init = 1;
}
/* Parameters the fork-only crate fixes; the twins of lsp_set_field_consts / lsp_set_transcript_flags. */
static int g_alpha_before_openings = 1, g_observe_opened_values = 0;
int lsp_oracle_set_field_consts(const uint64_t* generator, const uint64_t* two_adic_root_2_47) {
    init_consts();
    fr g, w; memcpy(&g, generator, 32); memcpy(&w, two_adic_root_2_47, 32);
    fr t = w; for (int i = 0; i < TWO_ADICITY - 1; i++) t = fr_sqr(t);
    if (fr_is_zero(g) || !fr_eq(t, fr_neg(ONE))) return -1;
    GEN = g; GEN_INV = fr_inv(g); ROOT47 = w;
    return 0;
}
void lsp_oracle_set_transcript_flags(int alpha_before_openings, int observe_opened_values) {
    g_alpha_before_openings = alpha_before_openings != 0; g_observe_opened_values = observe_opened_values != 0;
}
static fr two_adic_generator(int bits) { fr w = ROOT47; for (int i = bits; i < TWO_ADICITY; i++) w = fr_sqr(w); return w; }
static inline uint32_t bitrev(uint32_t x, int bits) {
    uint32_t r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r;
}
static void batch_inverse(fr* x, size_t n) { /* in place, all non-zero */
    if (n == 0) return;
    fr* pre = (fr*)malloc(n * sizeof(fr));
    fr acc = ONE;
    for (size_t i = 0; i < n; i++) { pre[i] = acc; acc = fr_mul(acc, x[i]); }
    acc = fr_inv(acc);
    for (size_t i = n; i-- > 0;) { fr t = fr_mul(acc, pre[i]); acc = fr_mul(acc, x[i]); x[i] = t; }
    free(pre);
}
static void par_batch_inverse(fr* x, size_t n) {
    size_t chunk = 4096;
    #pragma omp parallel for schedule(static)
    for (size_t s = 0; s < n; s += chunk) batch_inverse(x + s, (n - s < chunk) ? n - s : chunk);
}

/* ---- deterministic RNG (same as oracle/field.py SplitMix64) --------------- */
typedef struct { uint64_t s; } rng_t;
static uint64_t rng_u64(rng_t* g) {
    g->s += 0x9E3779B97F4A7C15ull; uint64_t z = g->s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
static fr rng_fr(rng_t* g) { /* canonical -> Montgomery */
    for (;;) {
        fr v; for (int i = 0; i < 4; i++) v.l[i] = rng_u64(g);
        v.l[3] &= (1ull << 61) - 1;
        if (!geq_p(v.l)) return fr_mul(v, R2);
    }
}

/* ---- Poseidon2 ------------------------------------------------ */
#define MAX_HALF_F 8
#define MAX_P 64
static struct { int d, half_f, rounds_p; fr ini[MAX_HALF_F][3], ter[MAX_HALF_F][3], in[MAX_P], diag[3]; int set; } PP;

int lsp_oracle_set_poseidon2(int sbox_d, int rounds_f, int rounds_p, const uint64_t* constants, const uint64_t* diag) {
    init_consts();
    if (rounds_f / 2 > MAX_HALF_F || rounds_p > MAX_P) return -1;
    PP.d = sbox_d; PP.half_f = rounds_f / 2; PP.rounds_p = rounds_p;
    const uint64_t* c = constants;
    for (int r = 0; r < PP.half_f; r++) for (int i = 0; i < 3; i++, c += 4) memcpy(&PP.ini[r][i], c, 32);
    for (int r = 0; r < PP.half_f; r++) for (int i = 0; i < 3; i++, c += 4) memcpy(&PP.ter[r][i], c, 32);
    for (int r = 0; r < rounds_p; r++, c += 4) memcpy(&PP.in[r], c, 32);
    memcpy(PP.diag, diag, 96);
    PP.set = 1;
    return 0;
}
static inline fr sbox(fr x) {
    fr x2 = fr_sqr(x);
    switch (PP.d) {
        case 3: return fr_mul(x2, x);
        case 5: return fr_mul(fr_sqr(x2), x);
        case 7: { fr x4 = fr_sqr(x2); return fr_mul(fr_mul(x4, x2), x); }
        case 11: { fr x8 = fr_sqr(fr_sqr(x2)); return fr_mul(fr_mul(x8, x2), x); }
        default: { fr x16 = fr_sqr(fr_sqr(fr_sqr(x2))); return fr_mul(x16, x); }
    }
}
static inline void ext_linear(fr* s) { fr t = fr_add(fr_add(s[0], s[1]), s[2]); s[0] = fr_add(s[0], t); s[1] = fr_add(s[1], t); s[2] = fr_add(s[2], t); }
static void permute(fr* s) {
    ext_linear(s);
    for (int r = 0; r < PP.half_f; r++) { for (int i = 0; i < 3; i++) s[i] = sbox(fr_add(s[i], PP.ini[r][i])); ext_linear(s); }
    for (int r = 0; r < PP.rounds_p; r++) {
        s[0] = sbox(fr_add(s[0], PP.in[r]));
        fr t = fr_add(fr_add(s[0], s[1]), s[2]);
        for (int i = 0; i < 3; i++) s[i] = fr_add(fr_mul(s[i], PP.diag[i]), t);
    }
    for (int r = 0; r < PP.half_f; r++) { for (int i = 0; i < 3; i++) s[i] = sbox(fr_add(s[i], PP.ter[r][i])); ext_linear(s); }
}
static fr hash_slice(const fr* x, size_t n) { /* PaddingFreeSponge<Perm,3,2,1>::hash_iter */
    fr s[3] = {ZERO, ZERO, ZERO};
    size_t i = 0;
    for (; i + 1 < n; i += 2) { s[0] = x[i]; s[1] = x[i + 1]; permute(s); }
    if (i < n) { s[0] = x[i]; permute(s); }
    return s[0];
}
static inline fr compress(fr l, fr r) { fr s[3] = {l, r, ZERO}; permute(s); return s[0]; }
void lsp_oracle_permute(const uint64_t* in, uint64_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) { fr s[3]; memcpy(s, in + 12 * i, 96); permute(s); memcpy(out + 12 * i, s, 96); }
}
void lsp_oracle_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) { fr x, y; memcpy(&x, a + 4 * i, 32); memcpy(&y, b + 4 * i, 32); x = fr_mul(x, y); memcpy(out + 4 * i, &x, 32); }
}

/* ---- Merkle tree (A.4); matrices are column-major: col c at base + c*rows ------- */
typedef struct { fr* dig; size_t h; int log_h; } tree_t;
static size_t layer_off(size_t h, int k) { return 2 * h - ((2 * h) >> k); }
static int ilog2(size_t n) { int k = 0; while (((size_t)1 << k) < n) k++; return k; }
static void tree_compress_up(tree_t* t) {
    for (int k = 0; k < t->log_h; k++) {
        const fr* in = t->dig + layer_off(t->h, k); fr* out = t->dig + layer_off(t->h, k + 1);
        size_t n = t->h >> (k + 1);
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) out[i] = compress(in[2 * i], in[2 * i + 1]);
    }
}
static tree_t tree_build_cols(const fr* const* cols, int width, size_t h) {
    tree_t t; t.h = h; t.log_h = ilog2(h); t.dig = (fr*)malloc((2 * h - 1) * sizeof(fr));
    #pragma omp parallel for schedule(static)
    for (size_t r = 0; r < h; r++) {
        fr row[256];
        for (int c = 0; c < width; c++) row[c] = cols[c][r];
        t.dig[r] = hash_slice(row, (size_t)width);
    }
    tree_compress_up(&t);
    return t;
}
static tree_t tree_build_pairs(const fr* v, size_t len) { /* rows (v[2j], v[2j+1]) */
    tree_t t; t.h = len / 2; t.log_h = ilog2(t.h); t.dig = (fr*)malloc((2 * t.h - 1) * sizeof(fr));
    #pragma omp parallel for schedule(static)
    for (size_t j = 0; j < t.h; j++) t.dig[j] = hash_slice(v + 2 * j, 2);
    tree_compress_up(&t);
    return t;
}
static fr tree_root(const tree_t* t) { return t->dig[2 * t->h - 2]; }
static fr tree_sibling(const tree_t* t, int k, size_t index) { return t->dig[layer_off(t->h, k) + ((index >> k) ^ 1)]; }

/* ---- NTT / LDE (A.3) ------------------------------------------------------------- */
static void ntt_inplace(fr* a, int log_n, fr root) { /* bit-reversed input -> natural output (DIT) */
    size_t n = (size_t)1 << log_n;
    for (int s = 0; s < log_n; s++) {
        size_t half = (size_t)1 << s;
        fr wm = root; for (int i = s + 1; i < log_n; i++) wm = fr_sqr(wm);
        for (size_t k = 0; k < n; k += 2 * half) {
            fr w = ONE;
            for (size_t j = 0; j < half; j++) {
                fr u = a[k + j], v = fr_mul(a[k + j + half], w);
                a[k + j] = fr_add(u, v); a[k + j + half] = fr_sub(u, v);
                w = fr_mul(w, wm);
            }
        }
    }
}
static void bitrev_permute(fr* a, int log_n) {
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) { size_t j = bitrev((uint32_t)i, log_n); if (i < j) { fr t = a[i]; a[i] = a[j]; a[j] = t; } }
}
/* coefficients of the interpolant of col over H_n, natural order */
static void idft_col(fr* col, int log_n) {
    size_t n = (size_t)1 << log_n;
    fr w = two_adic_generator(log_n);
    fr winv = fr_pow_u64(w, n - 1);
    bitrev_permute(col, log_n);
    ntt_inplace(col, log_n, winv);
    fr ninv = ONE; for (int i = 0; i < log_n; i++) ninv = fr_halve(ninv);
    for (size_t i = 0; i < n; i++) col[i] = fr_mul(col[i], ninv);
}
/* out (column-major, L rows per column, bit-reversed order) of in (column-major, n rows) */
static void coset_lde_cols(const fr* in, size_t n, int width, int added_bits, fr shift, fr* out, fr* coeffs_out) {
    int log_n = ilog2(n), log_l = log_n + added_bits; size_t big = (size_t)1 << log_l;
    fr wl = two_adic_generator(log_l);
    #pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < width; c++) {
        fr* co = (fr*)malloc(n * sizeof(fr));
        memcpy(co, in + (size_t)c * n, n * sizeof(fr));
        idft_col(co, log_n);
        if (coeffs_out) memcpy(coeffs_out + (size_t)c * n, co, n * sizeof(fr));
        fr* ev = out + (size_t)c * big;
        fr s = ONE;
        for (size_t i = 0; i < n; i++) { ev[i] = fr_mul(co[i], s); s = fr_mul(s, shift); }
        for (size_t i = n; i < big; i++) ev[i] = ZERO;
        /* ev[j] = p(shift w^j) natural, then stored bit-reversed: a DIT on bit-reversed input,
           followed by the output permutation */
        bitrev_permute(ev, log_l);
        ntt_inplace(ev, log_l, wl);
        bitrev_permute(ev, log_l);
        free(co);
    }
}

/* ---- AIR (air/src/lib.rs:116-167) ------------------------------------------------- */
typedef struct { uint32_t n_cols; const uint32_t* a_ids; const uint32_t* b_ids; uint32_t b_inverse_id, check_id; } air_cfg;

/* LogUp lookup configs (air/src/air_lookup.rs:2-11), registered with lsp_oracle_set_lookups before
 * prove / verify: `LineaAIR::eval` folds them BEFORE the permutation configs (trace/src/lib.rs:80-89). */
#define MAX_LK 8
#define MAX_LK_IDS 64
typedef struct { uint32_t n_a, n_t, n_b, a_filter, a_inv, check; uint32_t a_ids[MAX_LK_IDS]; uint32_t b_ids[MAX_LK_IDS];
                 uint32_t b_filter[MAX_LK_IDS], b_inv[MAX_LK_IDS], occ[MAX_LK_IDS]; } lookup_cfg;
static lookup_cfg LK[MAX_LK];
static int N_LK = 0;

/* blob per lookup: n_a, n_t, n_b, a_filter, a_inv, check, a_ids[n_a], then per table: b_filter, b_inv, occ, b_ids[n_b] */
int lsp_oracle_set_lookups(const uint32_t* blob, int n_lookups) {
    if (n_lookups < 0 || n_lookups > MAX_LK) return -1;
    const uint32_t* r = blob;
    for (int i = 0; i < n_lookups; i++) {
        lookup_cfg* c = &LK[i];
        c->n_a = r[0]; c->n_t = r[1]; c->n_b = r[2]; c->a_filter = r[3]; c->a_inv = r[4]; c->check = r[5];
        if (c->n_a > MAX_LK_IDS || c->n_t * c->n_b > MAX_LK_IDS || c->n_t > MAX_LK_IDS) return -1;
        r += 6;
        for (uint32_t j = 0; j < c->n_a; j++) c->a_ids[j] = *r++;
        for (uint32_t t = 0; t < c->n_t; t++) {
            c->b_filter[t] = *r++; c->b_inv[t] = *r++; c->occ[t] = *r++;
            for (uint32_t j = 0; j < c->n_b; j++) c->b_ids[t * c->n_b + j] = *r++;
        }
    }
    N_LK = n_lookups;
    return 0;
}
/* `get_log_quotient_degree`: a lookup's first-row constraint has degree 4 (=> 4 chunks), the permutation's 3 (=> 2) */
static int air_log_q(void) { return N_LK > 0 ? 2 : 1; }

/* eval_lookup (air/src/lib.rs:57-114), folded into acc */
static fr fold_lookup(const lookup_cfg* l, fr acc, const fr* local, const fr* next, fr alpha_air, fr delta,
                      fr is_first, fr is_last, fr is_trans, fr alpha) {
    fr a_l = ZERO;
    for (uint32_t j = 0; j < l->n_a; j++) a_l = fr_add(fr_mul(a_l, alpha_air), local[l->a_ids[j]]);
    a_l = fr_add(a_l, delta);
    acc = fr_add(fr_mul(acc, alpha), fr_sub(fr_mul(a_l, local[l->a_inv]), ONE));
    fr chk_l = fr_mul(local[l->a_filter], local[l->a_inv]), chk_n = fr_mul(next[l->a_filter], next[l->a_inv]);
    for (uint32_t t = 0; t < l->n_t; t++) {
        fr b_l = ZERO;
        for (uint32_t j = 0; j < l->n_b; j++) b_l = fr_add(fr_mul(b_l, alpha_air), local[l->b_ids[t * l->n_b + j]]);
        b_l = fr_add(b_l, delta);
        acc = fr_add(fr_mul(acc, alpha), fr_sub(fr_mul(b_l, local[l->b_inv[t]]), ONE));
        chk_l = fr_sub(chk_l, fr_mul(fr_mul(local[l->b_filter[t]], local[l->occ[t]]), local[l->b_inv[t]]));
        chk_n = fr_sub(chk_n, fr_mul(fr_mul(next[l->b_filter[t]], next[l->occ[t]]), next[l->b_inv[t]]));
    }
    acc = fr_add(fr_mul(acc, alpha), fr_mul(is_first, fr_sub(local[l->check], chk_l)));
    acc = fr_add(fr_mul(acc, alpha), fr_mul(is_trans, fr_sub(fr_sub(next[l->check], local[l->check]), chk_n)));
    acc = fr_add(fr_mul(acc, alpha), fr_mul(is_last, local[l->check]));
    return acc;
}

static fr fold_constraints(const air_cfg* cfgs, int n_cfgs, const fr* local, const fr* next, fr alpha_air, fr delta,
                           fr is_first, fr is_last, fr is_trans, fr alpha) {
    fr acc = ZERO;
    for (int k = 0; k < N_LK; k++) acc = fold_lookup(&LK[k], acc, local, next, alpha_air, delta, is_first, is_last, is_trans, alpha);
    for (int k = 0; k < n_cfgs; k++) {
        const air_cfg* p = &cfgs[k];
        fr a_l = ZERO, b_l = ZERO, a_n = ZERO;
        for (uint32_t j = 0; j < p->n_cols; j++) {
            a_l = fr_add(fr_mul(a_l, alpha_air), local[p->a_ids[j]]);
            b_l = fr_add(fr_mul(b_l, alpha_air), local[p->b_ids[j]]);
            a_n = fr_add(fr_mul(a_n, alpha_air), next[p->a_ids[j]]);
        }
        a_l = fr_add(a_l, delta); b_l = fr_add(b_l, delta); a_n = fr_add(a_n, delta);
        fr c0 = fr_sub(fr_mul(b_l, local[p->b_inverse_id]), ONE);
        fr c1 = fr_mul(is_first, fr_sub(local[p->check_id], fr_mul(a_l, local[p->b_inverse_id])));
        fr c2 = fr_mul(is_trans, fr_sub(next[p->check_id], fr_mul(fr_mul(local[p->check_id], a_n), next[p->b_inverse_id])));
        fr c3 = fr_mul(is_last, fr_sub(local[p->check_id], ONE));
        acc = fr_add(fr_mul(acc, alpha), c0); acc = fr_add(fr_mul(acc, alpha), c1);
        acc = fr_add(fr_mul(acc, alpha), c2); acc = fr_add(fr_mul(acc, alpha), c3);
    }
    return acc;
}

/* ---- HashChallenger<Val,Hash,1> (A.6) ---------------------------------------------- */
#define CH_CAP 4096
typedef struct { fr in[CH_CAP]; int n; } chal_t;
static void ch_observe(chal_t* c, fr x) { if (c->n < CH_CAP) c->in[c->n++] = x; }
static fr ch_sample(chal_t* c) { fr o = hash_slice(c->in, (size_t)c->n); c->in[0] = o; c->n = 1; return o; }
static uint64_t ch_sample_bits(chal_t* c, int bits) { fr v = fr_canonical(ch_sample(c)); return bits >= 64 ? v.l[0] : (v.l[0] & ((1ull << bits) - 1)); }
/* smallest witness (deterministic; the reference's rayon `find_any` is not): chunks of candidates tested in parallel */
static fr ch_grind(chal_t* c, int bits) {
    const uint64_t chunk = 4096;
    for (uint64_t base = 0;; base += chunk) {
        uint64_t best = UINT64_MAX;
        #pragma omp parallel for schedule(static) reduction(min : best)
        for (uint64_t w = base; w < base + chunk; w++) {
            if (w > best) continue;
            chal_t t; t.n = c->n; memcpy(t.in, c->in, (size_t)c->n * sizeof(fr));
            ch_observe(&t, fr_from_u64(w));
            if (ch_sample_bits(&t, bits) == 0 && w < best) best = w;
        }
        if (best != UINT64_MAX) { fr wf = fr_from_u64(best); ch_observe(c, wf); (void)ch_sample_bits(c, bits); return wf; }
    }
}

/* ---- prove ------------------------------------------------------------------------------- */
typedef struct { uint32_t log_blowup, log_final_poly_len, num_queries, proof_of_work_bits; } fri_cfg;

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

size_t lsp_oracle_proof_words(uint32_t log_n, uint32_t width, uint32_t log_q, const fri_cfg* fri) {
    size_t log_l = (size_t)log_n + fri->log_blowup, q = (size_t)1 << log_q, rounds = log_n - fri->log_final_poly_len;
    size_t f = (size_t)1 << (fri->log_blowup + fri->log_final_poly_len);
    size_t per_query = 1 + (width + log_l) + (q + log_l);
    for (size_t r = 0; r < rounds; r++) per_query += 1 + (log_l - 1 - r);
    return (2 + 2 * (size_t)width + q + rounds + f + 1 + (size_t)fri->num_queries * per_query) * 4;
}

/* barycentric interpolate_coset over the first n rows (coset g*H_n in bit-reversed order) of `cols` */
static void interpolate_coset_cols(const fr* const* cols, int width, int log_n, fr z, fr* ys) {
    size_t n = (size_t)1 << log_n;
    fr g = two_adic_generator(log_n);
    fr* den = (fr*)malloc(n * sizeof(fr)); fr* gp = (fr*)malloc(n * sizeof(fr));
    fr x = GEN, gpow = ONE;
    for (size_t i = 0; i < n; i++) { den[i] = fr_sub(z, x); gp[i] = gpow; x = fr_mul(x, g); gpow = fr_mul(gpow, g); }
    par_batch_inverse(den, n);
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) den[i] = fr_mul(den[i], gp[i]); /* col_scale, natural order */
    #pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < width; c++) {
        fr acc = ZERO;
        for (size_t i = 0; i < n; i++) acc = fr_add(acc, fr_mul(cols[c][bitrev((uint32_t)i, log_n)], den[i]));
        ys[c] = acc;
    }
    fr zn = z, sn = GEN; for (int i = 0; i < log_n; i++) { zn = fr_sqr(zn); sn = fr_sqr(sn); }
    fr zerofier = fr_sub(zn, sn);
    fr denom = fr_mul(fr_from_u64((uint64_t)n), fr_mul(sn, GEN_INV)); /* n * g^(n-1) */
    fr scale = fr_mul(zerofier, fr_inv(denom));
    for (int c = 0; c < width; c++) ys[c] = fr_mul(ys[c], scale);
    free(den); free(gp);
}

int lsp_oracle_prove(const fri_cfg* fri, const uint64_t* trace_rm, size_t rows, size_t width, const air_cfg* cfgs, int n_cfgs,
                     const uint64_t* publics, uint64_t* proof_out, size_t proof_words, double* timings /* 8 */) {
    init_consts();
    if (!PP.set) return -3;
    const size_t n = rows, W = width;
    const int log_n = ilog2(n), log_q = air_log_q(), q = 1 << log_q, log_b = (int)fri->log_blowup, log_l = log_n + log_b;
    if (((size_t)1 << log_n) != n || log_q > log_b || (int)fri->log_final_poly_len > log_n || log_b + (int)fri->log_final_poly_len > 10 || W > 250) return -1;
    const size_t L = (size_t)1 << log_l;
    const int n_rounds = log_n - (int)fri->log_final_poly_len, log_f = log_b + (int)fri->log_final_poly_len;
    size_t need = lsp_oracle_proof_words((uint32_t)log_n, (uint32_t)W, (uint32_t)log_q, fri);
    if (proof_words < need) return -1;
    fr* proof = (fr*)proof_out;
    fr *p_local = proof + 2, *p_next = p_local + W, *p_chunks = p_next + W, *p_commits = p_chunks + q,
       *p_final = p_commits + n_rounds, *p_pow = p_final + ((size_t)1 << log_f), *p_queries = p_pow + 1;
    fr pub[2]; memcpy(pub, publics, 64);
    double t0 = now_s(), t1;
    #define MARK(i) do { t1 = now_s(); if (timings) timings[i] = (t1 - t0) * 1e3; t0 = t1; } while (0)

    /* RowMajorMatrix -> column-major working copy */
    fr* tr = (fr*)malloc(n * W * sizeof(fr));
    const fr* rm = (const fr*)trace_rm;
    #pragma omp parallel for schedule(static)
    for (size_t r = 0; r < n; r++) for (size_t c = 0; c < W; c++) tr[c * n + r] = rm[r * W + c];

    /* commit to trace data */
    fr* lde_t = (fr*)malloc(L * W * sizeof(fr));
    coset_lde_cols(tr, n, (int)W, log_b, GEN, lde_t, NULL);
    MARK(0);
    const fr** cols_t = (const fr**)malloc(W * sizeof(fr*));
    for (size_t c = 0; c < W; c++) cols_t[c] = lde_t + c * L;
    tree_t tree_t_ = tree_build_cols(cols_t, (int)W, L);
    proof[0] = tree_root(&tree_t_);
    MARK(1);
    chal_t* ch = (chal_t*)malloc(sizeof(chal_t)); ch->n = 0;
    ch_observe(ch, fr_from_u64((uint64_t)log_n)); ch_observe(ch, proof[0]); ch_observe(ch, pub[0]); ch_observe(ch, pub[1]);
    fr alpha = ch_sample(ch);

    /* compute quotient polynomial: selectors by batch inversion, natural order over g*H_{Nq} */
    const int lnq = log_n + log_q; const size_t nq = (size_t)1 << lnq;
    fr* chunks = (fr*)malloc(nq * sizeof(fr)); /* q columns of n */
    {
        fr wnq = two_adic_generator(lnq), wn_inv = fr_pow_u64(two_adic_generator(log_n), n - 1);
        fr* xs = (fr*)malloc(nq * sizeof(fr)); fr* d1 = (fr*)malloc(nq * sizeof(fr)); fr* d2 = (fr*)malloc(nq * sizeof(fr));
        fr x = GEN; for (size_t i = 0; i < nq; i++) { xs[i] = x; x = fr_mul(x, wnq); }
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < nq; i++) { d1[i] = fr_sub(xs[i], ONE); d2[i] = fr_sub(xs[i], wn_inv); }
        par_batch_inverse(d1, nq); par_batch_inverse(d2, nq);
        fr gn = GEN; for (int i = 0; i < log_n; i++) gn = fr_sqr(gn);
        fr zh[8], zhi[8]; fr wq = two_adic_generator(log_q), wp = ONE;
        for (int c = 0; c < q; c++) { zh[c] = fr_sub(fr_mul(gn, wp), ONE); zhi[c] = fr_inv(zh[c]); wp = fr_mul(wp, wq); }
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < nq; i++) {
            fr local[256], next[256];
            size_t p = bitrev((uint32_t)i, lnq), pn = bitrev((uint32_t)((i + q) & (nq - 1)), lnq);
            for (size_t c = 0; c < W; c++) { local[c] = lde_t[c * L + p]; next[c] = lde_t[c * L + pn]; }
            int c = (int)(i & (q - 1));
            fr acc = fold_constraints(cfgs, n_cfgs, local, next, pub[0], pub[1], fr_mul(zh[c], d1[i]), fr_mul(zh[c], d2[i]),
                                      fr_sub(xs[i], wn_inv), alpha);
            chunks[(size_t)c * n + (i >> log_q)] = fr_mul(acc, zhi[c]);
        }
        free(xs); free(d1); free(d2);
    }
    MARK(2);

    /* commit to quotient poly chunks: shift_c = g / (g w_{Nq}^c) */
    fr* lde_q = (fr*)malloc((size_t)q * L * sizeof(fr));
    {
        fr wnq = two_adic_generator(lnq);
        for (int c = 0; c < q; c++) {
            fr shift = fr_pow_u64(wnq, (nq - (size_t)c) & (nq - 1));
            coset_lde_cols(chunks + (size_t)c * n, n, 1, log_b, shift, lde_q + (size_t)c * L, NULL);
        }
    }
    const fr* cols_q[8]; for (int c = 0; c < q; c++) cols_q[c] = lde_q + (size_t)c * L;
    tree_t tree_q = tree_build_cols(cols_q, q, L);
    proof[1] = tree_root(&tree_q);
    MARK(3);
    ch_observe(ch, proof[1]);
    fr zeta = ch_sample(ch);
    fr zeta_next = fr_mul(zeta, two_adic_generator(log_n));

    /* open: alpha first (fork-era; or after the observed openings, see g_alpha_before_openings), inverse denominators,
       barycentric openings, reduced rows */
    fr a_fri = ZERO;
    if (g_alpha_before_openings) a_fri = ch_sample(ch);
    fr* fold_all = (fr*)malloc(2 * L * sizeof(fr));
    {
        fr wl = two_adic_generator(log_l);
        fr* sub = (fr*)malloc(L * sizeof(fr)); fr* e0 = (fr*)malloc(L * sizeof(fr)); fr* e1 = (fr*)malloc(L * sizeof(fr));
        fr x = GEN; for (size_t i = 0; i < L; i++) { sub[bitrev((uint32_t)i, log_l)] = x; x = fr_mul(x, wl); }
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < L; i++) { e0[i] = fr_sub(sub[i], zeta); e1[i] = fr_sub(sub[i], zeta_next); }
        par_batch_inverse(e0, L); par_batch_inverse(e1, L);
        interpolate_coset_cols(cols_t, (int)W, log_n, zeta, p_local);
        interpolate_coset_cols(cols_t, (int)W, log_n, zeta_next, p_next);
        for (int c = 0; c < q; c++) interpolate_coset_cols(&cols_q[c], 1, log_n, zeta, p_chunks + c);
        if (g_observe_opened_values) for (size_t i = 0; i < 2 * W + (size_t)q; i++) ch_observe(ch, p_local[i]);
        if (!g_alpha_before_openings) a_fri = ch_sample(ch);
        fr ap[512]; ap[0] = ONE; for (size_t i = 1; i < 2 * W + q + 1; i++) ap[i] = fr_mul(ap[i - 1], a_fri);
        fr ry0 = ZERO, ry1 = ZERO; for (size_t i = 0; i < W; i++) { ry0 = fr_add(ry0, fr_mul(ap[i], p_local[i])); ry1 = fr_add(ry1, fr_mul(ap[i], p_next[i])); }
        #pragma omp parallel for schedule(static)
        for (size_t r = 0; r < L; r++) {
            fr rr = ZERO; for (size_t c = 0; c < W; c++) rr = fr_add(rr, fr_mul(ap[c], lde_t[c * L + r]));
            fr ro = fr_mul(fr_sub(rr, ry0), e0[r]);
            ro = fr_add(ro, fr_mul(fr_mul(ap[W], fr_sub(rr, ry1)), e1[r]));
            for (int c = 0; c < q; c++) ro = fr_add(ro, fr_mul(fr_mul(ap[2 * W + c], fr_sub(lde_q[(size_t)c * L + r], p_chunks[c])), e0[r]));
            fold_all[r] = ro;
        }
        free(sub); free(e0); free(e1);
    }
    MARK(4);

    /* FRI commit phase */
    tree_t* ftrees = (tree_t*)malloc((size_t)(n_rounds ? n_rounds : 1) * sizeof(tree_t));
    fr** fvecs = (fr**)malloc((size_t)(n_rounds + 1) * sizeof(fr*));
    {
        fr* cur = fold_all; size_t len = L;
        for (int r = 0; r < n_rounds; r++) {
            ftrees[r] = tree_build_pairs(cur, len); fvecs[r] = cur;
            p_commits[r] = tree_root(&ftrees[r]);
            ch_observe(ch, p_commits[r]);
            fr beta = ch_sample(ch);
            size_t h = len / 2; int log_h = ilog2(h);
            fr g_inv = fr_pow_u64(two_adic_generator(log_h + 1), len - 1);
            fr half_beta = fr_halve(beta);
            fr* pw = (fr*)malloc(h * sizeof(fr));
            fr acc = half_beta; for (size_t j = 0; j < h; j++) { pw[bitrev((uint32_t)j, log_h)] = acc; acc = fr_mul(acc, g_inv); }
            fr* nxt = cur + len;
            #pragma omp parallel for schedule(static)
            for (size_t j = 0; j < h; j++)
                nxt[j] = fr_add(fr_mul(fr_add(HALF, pw[j]), cur[2 * j]), fr_mul(fr_sub(HALF, pw[j]), cur[2 * j + 1]));
            free(pw); cur = nxt; len = h;
        }
        fvecs[n_rounds] = cur;
        fr fin[1024]; size_t f = (size_t)1 << log_f;
        for (size_t j = 0; j < f; j++) fin[j] = cur[bitrev((uint32_t)j, log_f)];
        idft_col(fin, log_f);
        for (size_t j = 0; j < f; j++) { p_final[j] = fin[j]; ch_observe(ch, fin[j]); }
    }
    MARK(5);
    *p_pow = ch_grind(ch, (int)fri->proof_of_work_bits);
    size_t per_query = (need / 4 - (size_t)(p_queries - proof)) / fri->num_queries;
    for (uint32_t qi = 0; qi < fri->num_queries; qi++) {
        size_t index = (size_t)ch_sample_bits(ch, log_l);
        fr* out = p_queries + (size_t)qi * per_query; size_t o = 0;
        fr iv = {{index, 0, 0, 0}}; out[o++] = iv;
        for (size_t c = 0; c < W; c++) out[o++] = lde_t[c * L + index];
        for (int k = 0; k < log_l; k++) out[o++] = tree_sibling(&tree_t_, k, index);
        for (int c = 0; c < q; c++) out[o++] = lde_q[(size_t)c * L + index];
        for (int k = 0; k < log_l; k++) out[o++] = tree_sibling(&tree_q, k, index);
        for (int r = 0; r < n_rounds; r++) {
            size_t ii = index >> r;
            out[o++] = fvecs[r][ii ^ 1];
            for (int k = 0; k < ftrees[r].log_h; k++) out[o++] = tree_sibling(&ftrees[r], k, ii >> 1);
        }
    }
    MARK(6);
    for (int r = 0; r < n_rounds; r++) free(ftrees[r].dig);
    free(ftrees); free(fvecs); free(fold_all); free(lde_q); free(tree_q.dig); free(chunks); free(ch);
    free(tree_t_.dig); free(cols_t); free(lde_t); free(tr);
    if (timings) timings[7] = 0;
    return 0;
}

/* ---- verify (A.11 + pcs/fri verifier) ---------------------------------------------------- */
static int verify_batch(fr root, int log_h, size_t index, const fr* row, size_t row_len, const fr* sib) {
    fr node = hash_slice(row, row_len);
    for (int k = 0; k < log_h; k++) node = ((index >> k) & 1) ? compress(sib[k], node) : compress(node, sib[k]);
    return fr_eq(node, root);
}
/* returns 0 = accepted, otherwise the failing check:
   1 shape, 2 trace opening, 3 quotient opening, 4 commit-phase opening, 5 final poly, 6 pow, 7 OOD */
int lsp_oracle_verify(const fri_cfg* fri, uint32_t log_n, size_t width, const air_cfg* cfgs, int n_cfgs, const uint64_t* publics,
                      const uint64_t* proof_in, size_t proof_words) {
    init_consts();
    if (!PP.set) return -3;
    const size_t W = width; const int log_q = air_log_q(), q = 1 << log_q, log_b = (int)fri->log_blowup, log_l = (int)log_n + log_b;
    const int n_rounds = (int)log_n - (int)fri->log_final_poly_len, log_f = log_b + (int)fri->log_final_poly_len;
    if (proof_words != lsp_oracle_proof_words(log_n, (uint32_t)W, (uint32_t)log_q, fri)) return 1;
    const fr* proof = (const fr*)proof_in;
    for (size_t i = 0; i < proof_words / 4; i++)      /* deserialisation: every element must be canonical (< r) */
        if (geq_p(proof[i].l)) return 1;
    const fr *p_local = proof + 2, *p_next = p_local + W, *p_chunks = p_next + W, *p_commits = p_chunks + q,
             *p_final = p_commits + n_rounds, *p_pow = p_final + ((size_t)1 << log_f), *p_queries = p_pow + 1;
    fr pub[2]; memcpy(pub, publics, 64);
    chal_t* ch = (chal_t*)malloc(sizeof(chal_t)); ch->n = 0;
    ch_observe(ch, fr_from_u64(log_n)); ch_observe(ch, proof[0]); ch_observe(ch, pub[0]); ch_observe(ch, pub[1]);
    fr alpha = ch_sample(ch);
    ch_observe(ch, proof[1]);
    fr zeta = ch_sample(ch), zeta_next = fr_mul(zeta, two_adic_generator((int)log_n));
    fr a_fri = ZERO;
    if (g_alpha_before_openings) a_fri = ch_sample(ch);
    if (g_observe_opened_values) for (size_t i = 0; i < 2 * width + (size_t)q; i++) ch_observe(ch, p_local[i]);
    if (!g_alpha_before_openings) a_fri = ch_sample(ch);
    fr betas[64];
    for (int r = 0; r < n_rounds; r++) { ch_observe(ch, p_commits[r]); betas[r] = ch_sample(ch); }
    for (size_t j = 0; j < ((size_t)1 << log_f); j++) ch_observe(ch, p_final[j]);
    ch_observe(ch, *p_pow);
    int rc = 0;
    if (ch_sample_bits(ch, (int)fri->proof_of_work_bits) != 0) rc = 6;
    size_t per_query = (proof_words / 4 - (size_t)(p_queries - proof)) / fri->num_queries;
    fr wl = two_adic_generator(log_l);
    for (uint32_t qi = 0; qi < fri->num_queries && !rc; qi++) {
        size_t index = (size_t)ch_sample_bits(ch, log_l);
        const fr* in = p_queries + (size_t)qi * per_query; size_t o = 0;
        if (in[o].l[0] != index || in[o].l[1] || in[o].l[2] || in[o].l[3]) { rc = 1; break; }   /* the stored index is a raw integer */
        o++;
        const fr* trow = in + o; o += W; const fr* tsib = in + o; o += log_l;
        const fr* qrow = in + o; o += q; const fr* qsib = in + o; o += log_l;
        if (!verify_batch(proof[0], log_l, index, trow, W, tsib)) { rc = 2; break; }
        if (!verify_batch(proof[1], log_l, index, qrow, (size_t)q, qsib)) { rc = 3; break; }
        fr x = fr_mul(GEN, fr_pow_u64(wl, bitrev((uint32_t)index, log_l)));
        fr ix0 = fr_inv(fr_sub(x, zeta)), ix1 = fr_inv(fr_sub(x, zeta_next));
        fr ap = ONE, ro = ZERO;
        for (size_t c = 0; c < W; c++) { ro = fr_add(ro, fr_mul(ap, fr_mul(fr_sub(trow[c], p_local[c]), ix0))); ap = fr_mul(ap, a_fri); }
        for (size_t c = 0; c < W; c++) { ro = fr_add(ro, fr_mul(ap, fr_mul(fr_sub(trow[c], p_next[c]), ix1))); ap = fr_mul(ap, a_fri); }
        for (int c = 0; c < q; c++) { ro = fr_add(ro, fr_mul(ap, fr_mul(fr_sub(qrow[c], p_chunks[c]), ix0))); ap = fr_mul(ap, a_fri); }
        fr folded = ZERO; size_t di = index;
        for (int r = 0; r < n_rounds; r++) {
            int log_fh = log_l - 1 - r;
            if (r == 0) folded = fr_add(folded, ro);
            fr ev[2] = {folded, folded};
            ev[(di ^ 1) & 1] = in[o]; o++;
            if (!verify_batch(p_commits[r], log_fh, di >> 1, ev, 2, in + o)) { rc = 4; break; }
            o += (size_t)log_fh;
            di >>= 1;
            /* fold_row */
            fr x0 = fr_pow_u64(two_adic_generator(log_fh + 1), bitrev((uint32_t)di, log_fh));
            fr x1 = fr_neg(x0);
            folded = fr_add(ev[0], fr_mul(fr_mul(fr_sub(betas[r], x0), fr_sub(ev[1], ev[0])), fr_inv(fr_sub(x1, x0))));
        }
        if (rc) break;
        fr xx = fr_pow_u64(wl, bitrev((uint32_t)di, log_l)), xp = ONE, e = ZERO;
        for (size_t j = 0; j < ((size_t)1 << log_f); j++) { e = fr_add(e, fr_mul(p_final[j], xp)); xp = fr_mul(xp, xx); }
        if (!fr_eq(e, folded)) rc = 5;
    }
    free(ch);
    if (rc) return rc;
    /* OOD check */
    fr g = GEN; int lnq = (int)log_n + log_q;
    fr wnq = two_adic_generator(lnq);
    fr shifts[8]; for (int i = 0; i < q; i++) shifts[i] = fr_mul(g, fr_pow_u64(wnq, (uint64_t)i));
    fr quotient = ZERO;
    for (int i = 0; i < q; i++) {
        fr zp = ONE;
        for (int j = 0; j < q; j++) if (j != i) {
            fr si = fr_inv(shifts[j]);
            fr a = fr_mul(zeta, si), b = fr_mul(shifts[i], si);
            for (unsigned k = 0; k < log_n; k++) { a = fr_sqr(a); b = fr_sqr(b); }
            zp = fr_mul(zp, fr_mul(fr_sub(a, ONE), fr_inv(fr_sub(b, ONE))));
        }
        quotient = fr_add(quotient, fr_mul(zp, p_chunks[i]));
    }
    fr zn = zeta; for (unsigned k = 0; k < log_n; k++) zn = fr_sqr(zn);
    fr z_h = fr_sub(zn, ONE);
    fr wn_inv = fr_pow_u64(two_adic_generator((int)log_n), ((uint64_t)1 << log_n) - 1);
    fr is_first = fr_mul(z_h, fr_inv(fr_sub(zeta, ONE))), is_last = fr_mul(z_h, fr_inv(fr_sub(zeta, wn_inv)));
    fr folded_c = fold_constraints(cfgs, n_cfgs, p_local, p_next, pub[0], pub[1], is_first, is_last, fr_sub(zeta, wn_inv), alpha);
    if (!fr_eq(fr_mul(folded_c, fr_inv(z_h)), quotient)) return 7;
    return 0;
}

/* ---- witness generation (trace/src/permutation.rs:24-93, trace/src/lib.rs:94-106) ----------- */
/* columns 2c (inverse) and 2c+1 (running product) of a row-major N x (2c+2) matrix whose a/b columns are filled */
static int witness_columns(fr* out, size_t n, uint32_t c, fr alpha, fr delta) {
    size_t W = 2 * (size_t)c + 2;
    /* per-row inversion, as the reference does (:70); the row loop is split so the
       inversions can use every core, the running product stays sequential (:72) */
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        fr ac = ZERO, bc = ZERO;
        for (uint32_t j = 0; j < c; j++) { ac = fr_add(fr_mul(ac, alpha), out[i * W + j]); bc = fr_add(fr_mul(bc, alpha), out[i * W + c + j]); }
        fr bi = fr_inv(fr_add(bc, delta));
        out[i * W + 2 * c] = bi;
        out[i * W + 2 * c + 1] = fr_mul(fr_add(ac, delta), bi);
    }
    fr prev = ONE;
    for (size_t i = 0; i < n; i++) { prev = fr_mul(prev, out[i * W + 2 * c + 1]); out[i * W + 2 * c + 1] = prev; }
    return fr_eq(prev, ONE) ? 0 : -2;                 /* permutation.rs:76-79 */
}

/* seed -> alpha, delta, c columns a, b = rows of a shuffled; out: row-major N x (2c+2), Montgomery */
int lsp_oracle_gen_trace(uint64_t seed, uint32_t c, uint32_t log_n, uint64_t* publics_out, uint64_t* trace_out) {
    init_consts();
    size_t n = (size_t)1 << log_n, W = 2 * (size_t)c + 2;
    rng_t g = {seed};
    fr alpha = rng_fr(&g), delta = rng_fr(&g);
    memcpy(publics_out, &alpha, 32); memcpy(publics_out + 4, &delta, 32);
    fr* out = (fr*)trace_out;
    for (uint32_t j = 0; j < c; j++) for (size_t i = 0; i < n; i++) out[i * W + j] = rng_fr(&g);
    size_t* perm = (size_t*)malloc(n * sizeof(size_t));
    for (size_t i = 0; i < n; i++) perm[i] = i;
    for (size_t i = n - 1; i > 0; i--) { size_t j = rng_u64(&g) % (i + 1); size_t t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
    for (size_t i = 0; i < n; i++) for (uint32_t j = 0; j < c; j++) out[i * W + c + j] = out[perm[i] * W + j];
    free(perm);
    return witness_columns(out, n, c, alpha, delta);
}

/* `RawPermutationTrace::get_trace` + `RawTrace::get_trace` on given input columns: ab_rm is row-major
   rows x 2c (a columns first), Montgomery limbs; publics = [alpha, delta]; out: row-major rows x (2c+2). */
int lsp_oracle_permutation_trace(const uint64_t* ab_rm, size_t rows, uint32_t c, const uint64_t* publics, uint64_t* trace_out) {
    init_consts();
    size_t W = 2 * (size_t)c + 2;
    fr alpha, delta;
    memcpy(&alpha, publics, 32); memcpy(&delta, publics + 4, 32);
    fr* out = (fr*)trace_out;
    const fr* in = (const fr*)ab_rm;
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < rows; i++) for (uint32_t j = 0; j < 2 * c; j++) out[i * W + j] = in[i * 2 * c + j];
    return witness_columns(out, rows, c, alpha, delta);
}

/* n <= 0: one thread per online processor, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1). */
int lsp_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n <= 0) n = omp_get_num_procs();
    omp_set_num_threads(n);
    return n;
#else
    (void)n;
    return 1;
#endif
}

int lsp_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
