// Proof serialisation (host only): the flat array `lsp_prove_*` writes <-> a self-describing byte stream that follows the
// FIELD ORDER of `p3_uni_stark::Proof` the way `bincode` (fixed-width little-endian integers, u64 length
// before every Vec) would write the derived `Serialize`:
//
//   header   "LSPP" u32 version(1) u32 log_blowup u32 log_final_poly_len u32 num_queries u32 proof_of_work_bits
//            u32 width u32 log_q                                   (what a reader needs besides degree_bits)
//   Proof    commitments { trace: [F;1], quotient_chunks: [F;1] }
//            opened_values { trace_local: Vec<F>, trace_next: Vec<F>, quotient_chunks: Vec<Vec<F>> }
//            opening_proof: FriProof { commit_phase_commits: Vec<[F;1]>,
//                                      query_proofs: Vec<QueryProof { input_proof: Vec<BatchOpening { opened_values: Vec<Vec<F>>,
//                                                                                              opening_proof: Vec<[F;1]> }>,
//                                                                     commit_phase_openings: Vec<{ sibling_value: F,
//                                                                                                  opening_proof: Vec<[F;1]> }> }>,
//                                      final_poly: Vec<F>, pow_witness: F }
//            degree_bits: u64
//
// F = the canonical integer as 32 little-endian bytes (ark-serialize's `CanonicalSerialize` for Fp256, which is what a
// serde impl over arkworks writes).  The reference never serialises a proof and the fork's serde impl for
// `Bls12_377Fr` is not available, so this is a documented format of this library, not a claim about the fork's bytes.
// The per-query index of the flat layout is not part of `Proof` (the verifier samples it): it is dropped on the way out,
// and on the way in its slot is filled with the all-ones marker LSP_INDEX_NOT_CARRIED, for which `lsp_verify_air`
// substitutes the index it samples itself (a flat proof that does carry indices is still checked against the samples).
#include <cstring>
#include <vector>

#include "../../include/lsp_b200.h"

namespace {

typedef unsigned __int128 u128;
const uint64_t P[4] = {0x0a11800000000001ull, 0x59aa76fed0000001ull, 0x60b44d1e5c37b001ull, 0x12ab655e9a2ca556ull};
const uint64_t R2[4] = {0x25d577bab861857bull, 0xcc2c27b58860591full, 0xa7cc008fe5dc8593ull, 0x011fdae7eff1c939ull};
const uint64_t NINV = 0x0a117fffffffffffull;  // -r^-1 mod 2^64

// Montgomery product a*b/2^256 mod r (CIOS, 4 x 64-bit limbs)
void mont_mul(const uint64_t* a, const uint64_t* b, uint64_t* out) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        const uint64_t m = t[0] * NINV;
        c = (u128)m * P[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * P[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    bool ge = t[4] != 0;
    if (!ge) {
        ge = true;
        for (int i = 3; i >= 0; i--) {
            if (t[i] != P[i]) {
                ge = t[i] > P[i];
                break;
            }
        }
    }
    if (ge) {
        u128 bw = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)t[i] - P[i] - bw;
            t[i] = (uint64_t)d;
            bw = (d >> 64) & 1;
        }
    }
    memcpy(out, t, 32);
}
bool canonical(const uint64_t* v) {
    for (int i = 3; i >= 0; i--) {
        if (v[i] != P[i]) return v[i] < P[i];
    }
    return false;
}

struct Shape {
    uint32_t log_n, width, log_q, log_l, q, rounds, n_final;
    lsp_fri_config fri;
};
bool make_shape(uint32_t log_n, uint32_t width, uint32_t log_q, const lsp_fri_config* fri, Shape& s) {
    if (!fri || lsp_proof_words(log_n, width, log_q, fri) == 0) return false;
    s = Shape{log_n, width, log_q, log_n + fri->log_blowup, 1u << log_q, log_n - fri->log_final_poly_len,
              1u << (fri->log_blowup + fri->log_final_poly_len), *fri};
    return true;
}

struct Writer {
    uint8_t* p;
    size_t cap, n = 0;
    void bytes(const void* src, size_t k) {
        if (p && n + k <= cap) memcpy(p + n, src, k);
        n += k;
    }
    void u32(uint32_t v) { bytes(&v, 4); }
    void u64(uint64_t v) { bytes(&v, 8); }
    void felt(const uint64_t* mont) {  // Montgomery limbs -> canonical little-endian bytes
        if (!p) {                          // size-only pass: nothing is read
            n += 32;
            return;
        }
        static const uint64_t one[4] = {1, 0, 0, 0};
        uint64_t c[4];
        mont_mul(mont, one, c);
        bytes(c, 32);
    }
    void felts(const uint64_t* mont, size_t k, bool len_prefix) {
        if (len_prefix) u64(k);
        for (size_t i = 0; i < k; i++) felt(mont ? mont + 4 * i : nullptr);
    }
};
struct Reader {
    const uint8_t* p;
    size_t len, n = 0;
    bool ok = true;
    bool take(void* dst, size_t k) {
        if (!ok || n + k > len) return ok = false;
        memcpy(dst, p + n, k);
        n += k;
        return true;
    }
    uint32_t u32() { uint32_t v = 0; take(&v, 4); return v; }
    uint64_t u64() { uint64_t v = 0; take(&v, 8); return v; }
    void expect_len(uint64_t k) { if (u64() != k) ok = false; }
    void felt(uint64_t* mont) {  // canonical bytes -> Montgomery limbs; a non-canonical value is malformed
        uint64_t c[4];
        if (!take(c, 32)) return;
        if (!canonical(c)) { ok = false; return; }
        mont_mul(c, R2, mont);
    }
    void felts(uint64_t* mont, size_t k, bool len_prefix) {
        if (len_prefix) expect_len(k);
        for (size_t i = 0; ok && i < k; i++) felt(mont + 4 * i);
    }
};

size_t write_proof(const Shape& s, const uint64_t* flat, uint8_t* out, size_t cap) {
    Writer w{out, cap};
    w.bytes("LSPP", 4);
    w.u32(1);
    w.u32(s.fri.log_blowup); w.u32(s.fri.log_final_poly_len); w.u32(s.fri.num_queries); w.u32(s.fri.proof_of_work_bits);
    w.u32(s.width); w.u32(s.log_q);
    size_t off = 0;   // (flat == nullptr: size-only pass, no pointer is formed)
    auto adv = [&](size_t k) { const uint64_t* r = flat ? flat + off : nullptr; off += 4 * k; return r; };
    w.felts(adv(1), 1, false);                                   // commitments.trace
    w.felts(adv(1), 1, false);                                   // commitments.quotient_chunks
    w.felts(adv(s.width), s.width, true);                        // opened_values.trace_local
    w.felts(adv(s.width), s.width, true);                        // opened_values.trace_next
    w.u64(s.q);                                                  // opened_values.quotient_chunks: Vec<Vec<F>>
    for (uint32_t c = 0; c < s.q; c++) w.felts(adv(1), 1, true);
    const uint64_t* commits = adv(s.rounds);
    const uint64_t* final_poly = adv(s.n_final);
    const uint64_t* pow = adv(1);
    w.felts(commits, s.rounds, true);                            // opening_proof.commit_phase_commits
    w.u64(s.fri.num_queries);                                    // opening_proof.query_proofs
    for (uint32_t qi = 0; qi < s.fri.num_queries; qi++) {
        adv(1);                                                  // the stored index: not part of `Proof`
        w.u64(2);                                                // input_proof: one BatchOpening per round (trace, quotient)
        w.u64(1);                                                //   trace round: one matrix
        w.felts(adv(s.width), s.width, true);
        w.felts(adv(s.log_l), s.log_l, true);
        w.u64(s.q);                                              //   quotient round: q matrices of width 1
        for (uint32_t c = 0; c < s.q; c++) w.felts(adv(1), 1, true);
        w.felts(adv(s.log_l), s.log_l, true);
        w.u64(s.rounds);                                         // commit_phase_openings
        for (uint32_t r = 0; r < s.rounds; r++) {
            w.felts(adv(1), 1, false);                           //   sibling_value
            w.felts(adv(s.log_l - 1 - r), s.log_l - 1 - r, true);
        }
    }
    w.felts(final_poly, s.n_final, true);                        // opening_proof.final_poly
    w.felts(pow, 1, false);                                      // opening_proof.pow_witness
    w.u64(s.log_n);                                              // degree_bits
    return w.n;
}

}  // namespace

extern "C" size_t lsp_proof_serialized_bytes(uint32_t log_n, uint32_t width, uint32_t log_q, const lsp_fri_config* fri) {
    Shape s;
    if (!make_shape(log_n, width, log_q, fri, s)) return 0;
    return write_proof(s, nullptr, nullptr, 0);   // size-only pass
}

extern "C" int lsp_proof_serialize(const uint64_t* proof, size_t proof_words, uint32_t log_n, uint32_t width, uint32_t log_q,
                                   const lsp_fri_config* fri, uint8_t* out, size_t out_cap, size_t* out_len) {
    Shape s;
    if (!proof || !out || !make_shape(log_n, width, log_q, fri, s) || proof_words != lsp_proof_words(log_n, width, log_q, fri)) return LSP_ERR_PARAM;
    const size_t n = write_proof(s, proof, out, out_cap);
    if (out_len) *out_len = n;
    return n <= out_cap ? LSP_OK : LSP_ERR_PARAM;
}

extern "C" int lsp_proof_deserialize(const uint8_t* bytes, size_t len, uint32_t* log_n_out, uint32_t* width_out, uint32_t* log_q_out,
                                     lsp_fri_config* fri_out, uint64_t* proof_out, size_t proof_words_cap, size_t* proof_words_out) {
    if (!bytes || !log_n_out || !width_out || !log_q_out || !fri_out) return LSP_ERR_PARAM;
    Reader r{bytes, len};
    char magic[4] = {0, 0, 0, 0};
    r.take(magic, 4);
    if (!r.ok || memcmp(magic, "LSPP", 4) != 0 || r.u32() != 1) return LSP_ERR_PARAM;
    lsp_fri_config fri;
    fri.log_blowup = r.u32(); fri.log_final_poly_len = r.u32(); fri.num_queries = r.u32(); fri.proof_of_work_bits = r.u32();
    const uint32_t width = r.u32(), log_q = r.u32();
    // the header comes from outside: hold it to the ranges the prover accepts (check_fri_config) BEFORE anything walks the
    // shape it describes -- a stream promising 2^32 queries must cost nothing to refuse
    if (!r.ok || len < 8 || fri.log_blowup > 16 || fri.log_final_poly_len > 31 || log_q > fri.log_blowup || fri.num_queries == 0 ||
        fri.num_queries > 4096 || fri.proof_of_work_bits > 32 || width == 0 || width > (1u << 20))
        return LSP_ERR_PARAM;
    uint64_t log_n64;
    memcpy(&log_n64, bytes + len - 8, 8);                        // degree_bits closes the stream: the shape depends on it
    if (log_n64 > 31) return LSP_ERR_PARAM;
    Shape s;
    if (!make_shape(uint32_t(log_n64), width, log_q, &fri, s)) return LSP_ERR_PARAM;
    const size_t words = lsp_proof_words(s.log_n, width, log_q, &fri);
    // the stream's length is a function of the shape: a header that promises anything else is malformed (and must not
    // make a caller allocate for it)
    if (write_proof(s, nullptr, nullptr, 0) != len) return LSP_ERR_PARAM;
    *log_n_out = s.log_n; *width_out = width; *log_q_out = log_q; *fri_out = fri;
    if (proof_words_out) *proof_words_out = words;
    if (!proof_out) return LSP_OK;                               // shape query
    if (proof_words_cap < words) return LSP_ERR_PARAM;
    memset(proof_out, 0, words * 8);
    uint64_t* f = proof_out;
    auto adv = [&](size_t k) { uint64_t* p = f; f += 4 * k; return p; };
    r.felts(adv(1), 1, false);
    r.felts(adv(1), 1, false);
    r.felts(adv(width), width, true);
    r.felts(adv(width), width, true);
    r.expect_len(s.q);
    for (uint32_t c = 0; c < s.q; c++) r.felts(adv(1), 1, true);
    uint64_t* commits = adv(s.rounds);
    uint64_t* final_poly = adv(s.n_final);
    uint64_t* pow = adv(1);
    r.felts(commits, s.rounds, true);
    r.expect_len(fri.num_queries);
    for (uint32_t qi = 0; r.ok && qi < fri.num_queries; qi++) {
        uint64_t* index_slot = adv(1);
        index_slot[0] = index_slot[1] = index_slot[2] = index_slot[3] = ~0ull;   // LSP_INDEX_NOT_CARRIED
        r.expect_len(2);
        r.expect_len(1);
        r.felts(adv(width), width, true);
        r.felts(adv(s.log_l), s.log_l, true);
        r.expect_len(s.q);
        for (uint32_t c = 0; c < s.q; c++) r.felts(adv(1), 1, true);
        r.felts(adv(s.log_l), s.log_l, true);
        r.expect_len(s.rounds);
        for (uint32_t rr = 0; rr < s.rounds; rr++) {
            r.felts(adv(1), 1, false);
            r.felts(adv(s.log_l - 1 - rr), s.log_l - 1 - rr, true);
        }
    }
    r.felts(final_poly, s.n_final, true);
    r.felts(pow, 1, false);
    if (r.u64() != s.log_n || !r.ok || r.n != len) return LSP_ERR_PARAM;
    return LSP_OK;
}
