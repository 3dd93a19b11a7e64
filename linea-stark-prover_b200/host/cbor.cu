// Input wire formats of the reference: `RawPermutationTrace` and `RawLookupTrace` as CBOR.
//
// Replaces `RawPermutationTrace::read_file` (reference trace/src/permutation.rs:17-22, ciborium +
// serde) and the zero-padding of `resize` (:134-142).  The struct is
//     { a: Vec<Vec<[u8;32]>>, b: Vec<Vec<[u8;32]>>, name: String }          (:9-14)
// and serde writes a `[u8;32]` as a 32-element CBOR ARRAY of small unsigned integers (not a byte
// string); byte strings of length 32 are accepted too.  Definite and indefinite lengths are both
// handled.  This file only parses: the 32-byte big-endian values are handed to the device as they
// are, and `from_be_bytes_mod_order` + the Montgomery conversion (:95-118) run in a kernel
// (lsp_permutation_trace_be, csrc/witness.cu) -- no field arithmetic happens on the host.
//
// Speed (the file is ~50 bytes of CBOR per element, 400 MB at 12 columns x 2^19
// rows, so a naive reader costs far more than the prove it feeds).  The structure pass is a serial
// scan -- element sizes vary, so a column's end is only known by walking it -- with a branch-light
// fast path for serde's canonical `98 20 <32 x (b | 18 b)>` element, and it records where every
// column starts; the decode pass then parses the columns on all host threads at once.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/lsp_b200.h"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define LSP_CBOR_SIMD 1
#endif

namespace {

#ifdef LSP_CBOR_SIMD
// Tables for the vectorised element decoder, indexed by an 8-bit mask of value-carrying bytes: the PSHUFB control that packs
// those bytes to the front of an 8-byte group, and the position of the k-th set bit.  (PEXT / PDEP would do both without
// tables, but they are microcoded -- hundreds of cycles -- on some hosts; the tables cost 4 KB of L1 and are fast everywhere.)
struct ElemTables {
    uint64_t pack[256];
    uint8_t kth[256][8];
    ElemTables() {
        for (unsigned m = 0; m < 256; m++) {
            uint64_t ctl = 0;
            unsigned k = 0;
            for (unsigned i = 0; i < 8; i++)
                if (m >> i & 1) {
                    ctl |= uint64_t(i) << (8 * k);
                    kth[m][k++] = uint8_t(i);
                }
            for (unsigned j = k; j < 8; j++) {
                ctl |= uint64_t(0x80) << (8 * j);    // PSHUFB writes zero
                kth[m][j] = 0;
            }
            pack[m] = ctl;
        }
    }
};
const ElemTables g_elem_tables;

// The canonical element, 64 value-region bytes at a time (AVX2; chosen at run time, the scalar loop below is the
// reference it is fuzzed against).  A value is one byte b < 0x18 or the pair `18 b`; which 0x18 bytes are MARKERS is a
// parity question inside runs of 0x18 (`18 18` is the value 24: marker, payload) -- the first byte of a run is a marker,
// then every second one.  With E = "byte == 0x18" and S = the run starts, adding the even-positioned starts to E clears
// exactly the runs that begin on an even position (the carry ripples through the run and lands behind it), which tells the
// two kinds of run apart; markers are the even positions of the first kind and the odd positions of the second.  Every
// non-marker byte carries a value; the 32nd of them ends the element (per-byte popcounts, a multiply for their prefix sums,
// one table look-up inside the group where the sum crosses 32); a non-marker byte that does not follow a marker must be
// < 0x18; PSHUFB packs the value bytes eight input bytes at a time.
__attribute__((target("avx2,popcnt"))) const uint8_t* fast_elem_simd(const uint8_t* p, uint8_t* dst) {
    // the caller guarantees 66 readable bytes at p
    if (p[0] != 0x98 || p[1] != 0x20) return nullptr;
    const uint8_t* q = p + 2;
    const __m256i lo = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(q));
    const __m256i hi = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(q + 32));
    const __m256i k18 = _mm256_set1_epi8(0x18), k19 = _mm256_set1_epi8(0x19);
    const uint64_t E = uint64_t(uint32_t(_mm256_movemask_epi8(_mm256_cmpeq_epi8(lo, k18)))) |
                       (uint64_t(uint32_t(_mm256_movemask_epi8(_mm256_cmpeq_epi8(hi, k18)))) << 32);
    const uint64_t G = uint64_t(uint32_t(_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_max_epu8(lo, k19), lo)))) |   // byte >= 0x19
                       (uint64_t(uint32_t(_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_max_epu8(hi, k19), hi)))) << 32);
    const uint64_t EVEN = 0x5555555555555555ull;
    const uint64_t S = E & ~(E << 1);
    const uint64_t r_even = E & ~(E + (S & EVEN));              // the bytes of runs that start on an even position
    const uint64_t M = (r_even & EVEN) | (E & ~r_even & ~EVEN);  // markers (never adjacent: at most 32 of the 64 bytes)
    const uint64_t V = ~M;                                      // value bytes: one-byte values and marker payloads
    uint64_t c = V - ((V >> 1) & EVEN);                         // popcount of every byte of V ...
    c = (c & 0x3333333333333333ull) + ((c >> 2) & 0x3333333333333333ull);
    c = (c + (c >> 4)) & 0x0f0f0f0f0f0f0f0full;
    const uint64_t pre = c * 0x0101010101010101ull;             // ... and their inclusive prefix sums (<= 64: no carries)
    const unsigned grp = unsigned(__builtin_ctzll(((pre | 0x8080808080808080ull) - 0x2020202020202020ull) & 0x8080808080808080ull)) >> 3;
    const unsigned before = grp ? unsigned(pre >> (8 * grp - 8)) & 0xff : 0;       // value bytes in the groups before it (< 32)
    const unsigned n = 8 * grp + g_elem_tables.kth[(V >> (8 * grp)) & 0xff][31 - before] + 1;   // up to and incl. the 32nd value byte
    const uint64_t used = n == 64 ? ~0ull : (1ull << n) - 1;
    if (G & V & ~(M << 1) & used) return nullptr;               // >= 0x19 where a one-byte value must stand: not this shape
    if (dst) {
        alignas(16) uint8_t tmp[48];
        const uint64_t vu = V & used;
        unsigned off = 0;
        for (unsigned g = 0; g <= grp; g++) {
            const unsigned m8 = unsigned(vu >> (8 * g)) & 0xff;
            const __m128i word = _mm_loadl_epi64(reinterpret_cast<const __m128i*>(q + 8 * g));
            const __m128i ctl = _mm_loadl_epi64(reinterpret_cast<const __m128i*>(&g_elem_tables.pack[m8]));
            _mm_storel_epi64(reinterpret_cast<__m128i*>(tmp + off), _mm_shuffle_epi8(word, ctl));
            off += unsigned(_mm_popcnt_u32(m8));
        }
        memcpy(dst, tmp, 32);
    }
    return q + n;
}
const bool g_cbor_simd = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("popcnt") && !getenv("LSP_CBOR_SCALAR");
#endif

// serde's canonical `[u8;32]`: array(32) head `98 20`, then 32 values, each one byte (< 24) or `18 b`.
// Returns the position after the element, or nullptr when the bytes at p are anything else.
inline const uint8_t* fast_elem(const uint8_t* p, const uint8_t* end, uint8_t* dst) {
    if (end - p < 34 || p[0] != 0x98 || p[1] != 0x20) return nullptr;
#ifdef LSP_CBOR_SIMD
    if (g_cbor_simd && end - p >= 66) return fast_elem_simd(p, dst);
#endif
    const uint8_t* q = p + 2;
    unsigned bad = 0;
    if (end - p >= 66) {            // the longest element fits: no bounds checks inside
        if (dst) {
            for (int k = 0; k < 32; k++) {
                const uint8_t b = q[0], wide = b == 0x18;
                bad |= b > 0x18;
                dst[k] = wide ? q[1] : b;
                q += 1 + wide;
            }
        } else {
            for (int k = 0; k < 32; k++) {
                const uint8_t b = q[0];
                bad |= b > 0x18;
                q += 1 + (b == 0x18);
            }
        }
        return bad ? nullptr : q;
    }
    for (int k = 0; k < 32; k++) {  // the last elements of the file
        if (q >= end) return nullptr;
        const uint8_t b = *q++;
        if (b > 0x18) return nullptr;
        uint8_t v = b;
        if (b == 0x18) {
            if (q >= end) return nullptr;
            v = *q++;
        }
        if (dst) dst[k] = v;
    }
    return q;
}

// A stretch of back-to-back canonical elements found by the parallel structure pass.  The serial
// structural walk jumps over it in one step and notes which vector it belongs to and at which row.
struct Run {
    size_t off, end, count;
    size_t vec_off = 0, row0 = 0;   // filled by the walk: head offset of the owning vector, first row
    bool placed = false;
};

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    const uint8_t* base = nullptr;   // start of the file, for the column offsets the structure pass records
    std::vector<Run>* runs = nullptr;   // sorted by off; when set, walk_vector jumps over them
    size_t run_hint = 0;                // runs before this index lie behind p
    size_t loose = 0;                   // elements met outside any run while runs are set

    // The run starting exactly at p, if any.
    Run* run_here() {
        if (!runs) return nullptr;
        const size_t o = size_t(p - base);
        while (run_hint < runs->size() && (*runs)[run_hint].off < o) run_hint++;
        if (run_hint < runs->size() && (*runs)[run_hint].off == o) return &(*runs)[run_hint];
        return nullptr;
    }

    bool need(size_t n) {
        if (size_t(end - p) < n) ok = false;
        return ok;
    }
    // Reads an item head; returns the major type, the argument in `arg`, `indef` for 0x1f.
    int head(uint64_t& arg, bool& indef) {
        indef = false;
        arg = 0;
        if (!need(1)) return -1;
        uint8_t b = *p++;
        int major = b >> 5, info = b & 31;
        if (info < 24) {
            arg = info;
        } else if (info <= 27) {
            int n = 1 << (info - 24);
            if (!need(size_t(n))) return -1;
            for (int i = 0; i < n; i++) arg = (arg << 8) | *p++;
        } else if (info == 31) {
            indef = true;
        } else {
            ok = false;
            return -1;
        }
        return major;
    }
    bool at_break() { return need(1) && *p == 0xff; }
    void skip_break() { p++; }

    // Skips one complete item of any type.  Nesting (arrays, maps, tags) is bounded like ciborium's recursion limit
    // (256): a longer run of 0x81 / 0xc0 / 0x9f bytes under an unknown key is a malformed file, not a stack overflow.
    static constexpr int MAX_DEPTH = 256;
    void skip(int depth = 0) {
        if (depth > MAX_DEPTH) {
            ok = false;
            return;
        }
        uint64_t arg;
        bool indef;
        int m = head(arg, indef);
        if (!ok) return;
        switch (m) {
            case 0: case 1: case 7: break;
            case 2: case 3:
                if (indef) {
                    while (ok && !at_break()) skip(depth + 1);
                    if (ok) skip_break();
                } else if (need(arg)) {
                    p += arg;
                }
                break;
            case 4: case 5: {
                uint64_t n = m == 5 ? 2 * arg : arg;
                if (indef) {
                    while (ok && !at_break()) skip(depth + 1);
                    if (ok) skip_break();
                } else {
                    for (uint64_t i = 0; ok && i < n; i++) skip(depth + 1);
                }
                break;
            }
            case 6: skip(depth + 1); break;
            default: ok = false;
        }
    }
    bool text(std::string& out) {
        uint64_t arg;
        bool indef;
        if (head(arg, indef) != 3 || indef || !need(arg)) return ok = false;
        out.assign(reinterpret_cast<const char*>(p), arg);
        p += arg;
        return true;
    }
    // One `[u8;32]`: array(32) of uints < 256, or bytes(32).  `dst` may be null (shape pass).
    bool elem32(uint8_t* dst) {
        if (const uint8_t* q = fast_elem(p, end, dst)) {
            p = q;
            return true;
        }
        uint64_t arg;
        bool indef;
        int m = head(arg, indef);
        if (!ok) return false;
        if (m == 2 && !indef && arg == 32 && need(32)) {
            if (dst) memcpy(dst, p, 32);
            p += 32;
            return true;
        }
        if (m != 4) return ok = false;
        int n = 0;
        while (ok && (indef ? !at_break() : uint64_t(n) < arg)) {
            uint64_t v;
            bool vi;
            if (head(v, vi) != 0 || vi || v > 255 || n >= 32) return ok = false;
            if (dst) dst[n] = uint8_t(v);
            n++;
        }
        if (ok && indef) skip_break();
        return ok = ok && n == 32;
    }
};

struct Shape {
    std::vector<size_t> a_rows, b_rows;  // rows per column
    std::vector<size_t> a_off, b_off;    // byte offset of every column's array head
    std::string name;
    size_t height() const {
        size_t h = 0;
        for (size_t r : a_rows) h = r > h ? r : h;
        for (size_t r : b_rows) h = r > h ? r : h;
        return h;
    }
};

// array(rows) of [u8;32]; row i goes to out + (i*stride + col)*32 when out != null and i < max_rows.  Returns the length.
size_t walk_vector(Reader& r, uint8_t* out, size_t stride, size_t col, size_t max_rows) {
    const size_t head_off = r.base ? size_t(r.p - r.base) : 0;   // (structure pass) identifies this vector: offset of its array head
    uint64_t nr;
    bool ir;
    if (r.head(nr, ir) != 4) {
        r.ok = false;
        return 0;
    }
    size_t i = 0;
    while (r.ok && (ir ? !r.at_break() : i < nr)) {
        if (Run* run = r.run_here()) {          // structure pass over a pre-scanned file: a whole run at once
            if (!ir && i + run->count > nr) {   // the run reaches past this vector: not the regular layout
                r.ok = false;
                break;
            }
            run->vec_off = head_off;
            run->row0 = i;
            run->placed = true;
            i += run->count;
            r.p = r.base + run->end;
            continue;
        }
        if (r.runs) r.loose++;
        r.elem32(out && i < max_rows ? out + (i * stride + col) * 32 : nullptr);
        i++;
    }
    if (r.ok && ir) r.skip_break();
    return i;
}

size_t host_threads() {
    size_t nt = std::thread::hardware_concurrency();
    if (const char* e = getenv("LSP_CBOR_THREADS")) nt = size_t(atoi(e));
    if (nt == 0) nt = 1;
    return nt > 64 ? 64 : nt;
}

// body(k) for k < n on up to host_threads() threads (dynamic schedule).
template <class F>
void parallel_for(size_t n, F body) {
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= n) break;
            body(k);
        }
    };
    size_t nt = host_threads();
    if (nt > n) nt = n;
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
}

// Parallel structure pre-pass.  In a file written by serde every element starts with `98 20`, and that byte pair
// cannot occur anywhere else (a payload byte 98 is always followed by a value head, never by 20), so a
// thread dropped at an arbitrary offset can resynchronise on it.  Each thread walks from its first marker to the
// next thread's, counting back-to-back canonical elements into runs and stepping over the structural tokens
// between them (array / map heads, breaks, text keys); it must land exactly on the next thread's marker.  Anything
// else -- another element encoding, a 32-row column (whose head is also `98 20`), a marker inside a string --
// makes the file "irregular": the caller falls back to the serial pass, which is the authority on malformed input.
bool prescan(const uint8_t* cbor, size_t len, std::vector<Run>& runs) {
    const size_t nt = host_threads();
    size_t min_len = size_t(4) << 20, chunk = size_t(1) << 18;   // below 4 MB the serial pass is quick enough
    if (const char* e = getenv("LSP_CBOR_PRESCAN_MIN")) {         // tests: force the parallel path on small files
        min_len = size_t(atoll(e));
        chunk = 64;
    }
    if (nt < 2 || len < min_len || len < 2 * chunk) return false;
    size_t n_chunks = len / chunk;
    if (n_chunks > 4 * nt) n_chunks = 4 * nt;
    const size_t npos = ~size_t(0);
    std::vector<size_t> first(n_chunks, npos);
    parallel_for(n_chunks, [&](size_t t) {
        const size_t s0 = len * t / n_chunks, e0 = len * (t + 1) / n_chunks;
        for (const uint8_t* q = cbor + s0; q < cbor + e0;) {
            q = static_cast<const uint8_t*>(memchr(q, 0x98, size_t(cbor + e0 - q)));
            if (!q) break;
            if (q + 1 < cbor + len && q[1] == 0x20) {
                first[t] = size_t(q - cbor);
                break;
            }
            q++;
        }
    });
    std::vector<size_t> bounds;
    for (size_t f : first)
        if (f != npos) bounds.push_back(f);
    if (bounds.empty()) return false;
    std::vector<std::vector<Run>> per(bounds.size());
    std::atomic<bool> regular{true};
    parallel_for(bounds.size(), [&](size_t k) {
        const uint8_t* end = cbor + len;
        const uint8_t* p = cbor + bounds[k];
        const uint8_t* lim = k + 1 < bounds.size() ? cbor + bounds[k + 1] : end;
        Run cur{0, 0, 0};
        while (p < lim && regular.load(std::memory_order_relaxed)) {
            if (const uint8_t* q = fast_elem(p, end, nullptr)) {
                if (cur.count == 0) cur.off = size_t(p - cbor);
                cur.count++;
                cur.end = size_t(q - cbor);
                p = q;
                continue;
            }
            if (cur.count) {
                per[k].push_back(cur);
                cur = Run{0, 0, 0};
            }
            Reader r{p, end};
            uint64_t arg;
            bool indef;
            const int m = r.head(arg, indef);
            bool fine = r.ok;
            if (fine && m == 4) fine = indef || arg != 32;          // array(32): an element in another form, or a 32-row column
            else if (fine && m == 3 && !indef) {                    // a map key or the name
                fine = r.need(arg);
                if (fine) r.p += arg;
            } else if (fine) fine = m == 5 || (m == 7 && indef);    // map head, break
            if (!fine) {
                regular = false;
                return;
            }
            p = r.p;
        }
        if (p != lim) regular = false;
        if (cur.count) per[k].push_back(cur);
    });
    if (getenv("LSP_CBOR_DEBUG")) fprintf(stderr, "prescan: %zu chunks, %zu bounds, regular=%d\n", n_chunks, bounds.size(), int(regular.load()));
    if (!regular) return false;
    runs.clear();
    for (auto& v : per) runs.insert(runs.end(), v.begin(), v.end());
    return !runs.empty();
}

// Zeroes rows [from, rows) of one output column (`resize` pads short columns with zeros).  Full-length columns --
// the normal case -- cost nothing: there is no blanket memset of the output, every cell is written exactly once,
// by the decoder or by the padding.
void pad_column(uint8_t* out, size_t stride, size_t col, size_t from, size_t rows) {
    for (size_t i = from; i < rows; i++) memset(out + (i * stride + col) * 32, 0, 32);
}

// One vector of the file (found by the structure pass) and the output column it fills.
struct VecJob {
    size_t off, col;
};

// Decode pass: every vector is parsed independently from its recorded offset, on all host threads.
bool decode_jobs(const uint8_t* cbor, size_t len, const std::vector<VecJob>& jobs, uint8_t* out, size_t stride, size_t max_rows) {
    std::atomic<bool> ok{true};
    parallel_for(jobs.size(), [&](size_t k) {
        Reader r{cbor + jobs[k].off, cbor + len};
        walk_vector(r, out, stride, jobs[k].col, max_rows);
        if (!r.ok) ok = false;
    });
    return ok;
}

// Decode pass over a pre-scanned file: one task per run (a slice of one column), so the work spreads over all
// threads whatever the number of columns.  Runs of vectors that are not in `jobs` (filters of absent tables) are dropped.
bool decode_runs(const uint8_t* cbor, size_t len, const std::vector<Run>& runs, const std::vector<VecJob>& jobs, uint8_t* out,
                 size_t stride, size_t max_rows) {
    std::vector<VecJob> by_off(jobs);
    std::sort(by_off.begin(), by_off.end(), [](const VecJob& a, const VecJob& b) { return a.off < b.off; });
    std::atomic<bool> ok{true};
    parallel_for(runs.size(), [&](size_t k) {
        const Run& run = runs[k];
        auto it = std::lower_bound(by_off.begin(), by_off.end(), run.vec_off, [](const VecJob& j, size_t o) { return j.off < o; });
        if (it == by_off.end() || it->off != run.vec_off) return;
        const uint8_t* p = cbor + run.off;
        for (size_t i = 0; i < run.count; i++) {
            const size_t row = run.row0 + i;
            p = fast_elem(p, cbor + len, row < max_rows ? out + (row * stride + it->col) * 32 : nullptr);
            if (!p) {
                ok = false;
                return;
            }
        }
    });
    return ok;
}

// After a structural walk over a pre-scanned file: every element was inside a run and every run was claimed by a vector.
bool runs_consumed(const Reader& r) {
    if (!r.runs) return true;
    if (r.loose) return false;
    for (const Run& run : *r.runs)
        if (!run.placed) return false;
    return true;
}

// Structure pass over `a` or `b`: array(columns) of array(rows) of [u8;32]; records every column's length and offset.
bool scan_columns(Reader& r, std::vector<size_t>& rows_out, std::vector<size_t>& off_out) {
    uint64_t nc;
    bool ic;
    if (r.head(nc, ic) != 4) return r.ok = false;
    size_t j = 0;
    while (r.ok && (ic ? !r.at_break() : j < nc)) {
        off_out.push_back(size_t(r.p - r.base));
        rows_out.push_back(walk_vector(r, nullptr, 0, 0, 0));
        j++;
    }
    if (r.ok && ic) r.skip_break();
    return r.ok;
}

// Structure pass over the top-level map of a RawPermutationTrace.
bool scan_impl(const uint8_t* cbor, size_t len, Shape& shape, std::vector<Run>* runs) {
    Reader r{cbor, cbor + len};
    r.base = cbor;
    r.runs = runs;
    uint64_t n;
    bool indef;
    if (r.head(n, indef) != 5) return false;
    bool seen_a = false, seen_b = false;
    size_t k = 0;
    while (r.ok && (indef ? !r.at_break() : k < n)) {
        std::string key;
        if (!r.text(key)) return false;
        if (key == "a") {
            if (seen_a) return false;                       // serde's derived Deserialize: "duplicate field"
            scan_columns(r, shape.a_rows, shape.a_off);
            seen_a = true;
        } else if (key == "b") {
            if (seen_b) return false;
            scan_columns(r, shape.b_rows, shape.b_off);
            seen_b = true;
        } else if (key == "name") {
            r.text(shape.name);
        } else {
            r.skip();
        }
        k++;
    }
    return r.ok && seen_a && seen_b && runs_consumed(r);
}

// Structure pass; with `runs` set it first tries the parallel pre-pass and leaves the placed runs there
// (empty when the file is not in the regular layout and the serial pass did the work).
bool scan(const uint8_t* cbor, size_t len, Shape& shape, std::vector<Run>* runs) {
    if (runs) {
        if (prescan(cbor, len, *runs)) {
            Shape s;
            if (scan_impl(cbor, len, s, runs)) {
                shape = std::move(s);
                return true;
            }
            if (getenv("LSP_CBOR_DEBUG")) fprintf(stderr, "scan: walk over %zu runs failed\n", runs->size());
        }
        runs->clear();
    }
    return scan_impl(cbor, len, shape, nullptr);
}

// ---- RawLookupTrace (trace/src/lookup.rs:10-17) -----------------------------------------------------
//   { a: Vec<Vec<[u8;32]>>, b: Vec<Vec<Vec<[u8;32]>>>, name, a_filter: Vec<[u8;32]>, b_filter: Vec<Vec<[u8;32]>> }
struct LookupShape {
    std::vector<size_t> a_rows, a_off;                  // per a column
    std::vector<std::vector<size_t>> b_rows, b_off;     // per table, per column
    size_t a_filter_len = 0, a_filter_off = 0;
    bool has_a_filter = false;
    std::vector<size_t> b_filter_len, b_filter_off;     // per table present in the file
    std::string name;
    size_t height() const {                     // get_max_height (:215-228)
        size_t h = 0;
        for (size_t r : a_rows) h = r > h ? r : h;
        for (auto& t : b_rows)
            for (size_t r : t) h = r > h ? r : h;
        return h;
    }
};

// Structure pass over the top-level map of a RawLookupTrace: lengths and offsets of every vector.
bool scan_lookup_impl(const uint8_t* cbor, size_t len, LookupShape& shape, std::vector<Run>* runs) {
    Reader r{cbor, cbor + len};
    r.base = cbor;
    r.runs = runs;
    uint64_t n;
    bool indef;
    if (r.head(n, indef) != 5) return false;
    bool seen_a = false, seen_b = false, seen_bf = false;
    size_t k = 0;
    while (r.ok && (indef ? !r.at_break() : k < n)) {
        std::string key;
        if (!r.text(key)) return false;
        uint64_t cnt;
        bool ic;
        if (key == "a") {
            if (seen_a) return false;
            scan_columns(r, shape.a_rows, shape.a_off);
            seen_a = true;
        } else if (key == "b") {
            if (seen_b) return false;
            if (r.head(cnt, ic) != 4) return false;
            size_t t = 0;
            while (r.ok && (ic ? !r.at_break() : t < cnt)) {
                shape.b_rows.emplace_back();
                shape.b_off.emplace_back();
                scan_columns(r, shape.b_rows.back(), shape.b_off.back());
                t++;
            }
            if (r.ok && ic) r.skip_break();
            seen_b = true;
        } else if (key == "a_filter") {
            if (shape.has_a_filter) return false;
            shape.a_filter_off = size_t(r.p - r.base);
            shape.a_filter_len = walk_vector(r, nullptr, 0, 0, 0);
            shape.has_a_filter = true;
        } else if (key == "b_filter") {
            if (seen_bf) return false;
            scan_columns(r, shape.b_filter_len, shape.b_filter_off);
            seen_bf = true;
        } else if (key == "name") {
            r.text(shape.name);
        } else {
            r.skip();
        }
        k++;
    }
    return r.ok && seen_a && seen_b && runs_consumed(r);
}

bool scan_lookup(const uint8_t* cbor, size_t len, LookupShape& shape, std::vector<Run>* runs) {
    if (runs) {
        if (prescan(cbor, len, *runs)) {
            LookupShape s;
            if (scan_lookup_impl(cbor, len, s, runs)) {
                shape = std::move(s);
                return true;
            }
        }
        runs->clear();
    }
    return scan_lookup_impl(cbor, len, shape, nullptr);
}

bool lookup_shape(const uint8_t* cbor, size_t len, LookupShape& s, std::vector<Run>* runs) {
    if (!scan_lookup(cbor, len, s, runs)) return false;
    if (s.a_rows.empty() || s.b_rows.empty() || s.b_rows[0].empty()) return false;
    for (auto& t : s.b_rows)
        if (t.size() != s.b_rows[0].size()) return false;   // AirLookupConfig::width assumes equal table widths (air_lookup.rs:37-39)
    return true;
}

// Fills the row-major output of a scanned RawLookupTrace (values, zero padding, default filters).
bool fill_lookup(const uint8_t* cbor, size_t len, const LookupShape& s, const std::vector<Run>& runs, uint8_t* out, size_t rows) {
    const size_t n_a = s.a_rows.size(), n_t = s.b_rows.size(), n_b = s.b_rows[0].size();
    const size_t stride = n_a + n_t * n_b + 1 + n_t, col_af = n_a + n_t * n_b;
    // `resize` pads columns and filters with zeros (:230-246)
    for (size_t j = 0; j < n_a; j++) pad_column(out, stride, j, s.a_rows[j], rows);
    for (size_t t = 0; t < n_t; t++)
        for (size_t j = 0; j < n_b; j++) pad_column(out, stride, n_a + t * n_b + j, s.b_rows[t][j], rows);
    pad_column(out, stride, col_af, s.has_a_filter ? s.a_filter_len : 0, rows);
    for (size_t t = 0; t < n_t; t++) pad_column(out, stride, col_af + 1 + t, t < s.b_filter_len.size() ? s.b_filter_len[t] : 0, rows);
    std::vector<VecJob> jobs;
    for (size_t j = 0; j < n_a; j++) jobs.push_back({s.a_off[j], j});
    for (size_t t = 0; t < n_t; t++)
        for (size_t j = 0; j < n_b; j++) jobs.push_back({s.b_off[t][j], n_a + t * n_b + j});
    if (s.has_a_filter) jobs.push_back({s.a_filter_off, col_af});
    for (size_t t = 0; t < s.b_filter_off.size() && t < n_t; t++)     // filters of tables the file does not have are dropped
        jobs.push_back({s.b_filter_off[t], col_af + 1 + t});
    if (!(runs.empty() ? decode_jobs(cbor, len, jobs, out, stride, rows) : decode_runs(cbor, len, runs, jobs, out, stride, rows))) return false;
    // `read_file` (:25-41): missing filter entries default to ONE up to the length of the first column they guard
    auto fill_ones = [&](size_t col, size_t from, size_t to) {
        for (size_t i = from; i < to && i < rows; i++) out[(i * stride + col) * 32 + 31] = 1;
    };
    fill_ones(col_af, s.a_filter_len, s.a_rows[0]);
    for (size_t t = 0; t < n_t; t++) fill_ones(col_af + 1 + t, t < s.b_filter_len.size() ? s.b_filter_len[t] : 0, s.b_rows[t][0]);
    return true;
}

// The same for a scanned RawPermutationTrace.
bool fill_permutation(const uint8_t* cbor, size_t len, const Shape& s, const std::vector<Run>& runs, uint8_t* out, size_t rows) {
    const size_t nc = s.a_rows.size();
    std::vector<VecJob> jobs;
    for (size_t j = 0; j < nc; j++) {                        // short columns are zero-padded (`resize`, permutation.rs:134-142)
        pad_column(out, 2 * nc, j, s.a_rows[j], rows);
        pad_column(out, 2 * nc, nc + j, s.b_rows[j], rows);
        jobs.push_back({s.a_off[j], j});
        jobs.push_back({s.b_off[j], nc + j});
    }
    return runs.empty() ? decode_jobs(cbor, len, jobs, out, 2 * nc, rows) : decode_runs(cbor, len, runs, jobs, out, 2 * nc, rows);
}

bool permutation_shape(const uint8_t* cbor, size_t len, Shape& s, std::vector<Run>* runs) {
    if (!scan(cbor, len, s, runs)) return false;
    return !s.a_rows.empty() && s.a_rows.size() == s.b_rows.size();   // air/src/lib.rs zips a and b ids
}

void copy_name(const std::string& from, char* name, size_t name_cap) {
    if (!name || !name_cap) return;
    const size_t n = from.size() < name_cap - 1 ? from.size() : name_cap - 1;
    memcpy(name, from.data(), n);
    name[n] = 0;
}

// ---- output buffers of the `_read` entry points ------------------------------------------------------------
// Plain malloc by default.  After lsp_host_pinned(1) they are page-locked (cudaHostAllocPortable: DMA-able from every
// device of the process), so the upload that follows runs at PCIe rate instead of through the driver's staging path,
// and eight ranks can read the same buffer at once.  Pinning costs ~0.4 ms per MB, so freed blocks are kept and
// handed out again (best fit): a prover that reads file after file pins once.
struct PinnedPool {
    std::mutex mu;
    bool enabled = false;
    std::map<void*, size_t> live;                 // handed out
    std::multimap<size_t, void*> spare;           // returned, by size
} g_pool;

void* host_alloc(size_t n) {
    if (n == 0) n = 1;
    {
        std::lock_guard<std::mutex> lock(g_pool.mu);
        if (!g_pool.enabled) return malloc(n);
        auto it = g_pool.spare.lower_bound(n);
        if (it != g_pool.spare.end()) {
            void* p = it->second;
            g_pool.live[p] = it->first;
            g_pool.spare.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, n, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();                       // no device / no lockable memory: an ordinary buffer still works
        return malloc(n);
    }
    std::lock_guard<std::mutex> lock(g_pool.mu);
    g_pool.live[p] = n;
    return p;
}

void host_release(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(g_pool.mu);
        auto it = g_pool.live.find(p);
        if (it != g_pool.live.end()) {
            g_pool.spare.emplace(it->second, p);
            g_pool.live.erase(it);
            return;
        }
    }
    free(p);
}

}  // namespace

extern "C" int lsp_host_pinned(int enable) {
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> lock(g_pool.mu);
        g_pool.enabled = enable != 0;
        if (!g_pool.enabled) {
            for (auto& kv : g_pool.spare) drop.push_back(kv.second);
            g_pool.spare.clear();
        }
    }
    for (void* p : drop) cudaFreeHost(p);
    return LSP_OK;
}

extern "C" int lsp_cbor_lookup_shape(const uint8_t* cbor, size_t len, size_t* rows, uint32_t* n_a_cols, uint32_t* n_tables,
                                     uint32_t* n_b_cols, char* name, size_t name_cap) {
    if (!cbor || !rows || !n_a_cols || !n_tables || !n_b_cols) return LSP_ERR_PARAM;
    LookupShape s;
    std::vector<Run> runs;
    if (!lookup_shape(cbor, len, s, &runs)) return LSP_ERR_PARAM;
    *rows = s.height();
    *n_a_cols = uint32_t(s.a_rows.size());
    *n_tables = uint32_t(s.b_rows.size());
    *n_b_cols = uint32_t(s.b_rows[0].size());
    copy_name(s.name, name, name_cap);
    return LSP_OK;
}

extern "C" int lsp_cbor_lookup_decode(const uint8_t* cbor, size_t len, uint8_t* be_rowmajor, size_t rows, uint32_t n_a_cols,
                                      uint32_t n_tables, uint32_t n_b_cols) {
    if (!cbor || !be_rowmajor || rows == 0) return LSP_ERR_PARAM;
    LookupShape s;
    std::vector<Run> runs;
    if (!lookup_shape(cbor, len, s, &runs)) return LSP_ERR_PARAM;
    if (s.height() > rows || s.a_rows.size() != n_a_cols || s.b_rows.size() != n_tables || s.b_rows[0].size() != n_b_cols) return LSP_ERR_PARAM;
    return fill_lookup(cbor, len, s, runs, be_rowmajor, rows) ? LSP_OK : LSP_ERR_PARAM;
}

extern "C" int lsp_cbor_lookup_read(const uint8_t* cbor, size_t len, size_t* rows, uint32_t* n_a_cols, uint32_t* n_tables,
                                    uint32_t* n_b_cols, char* name, size_t name_cap, uint8_t** be_rowmajor_out) {
    return lsp_cbor_lookup_read_rows(cbor, len, 0, rows, n_a_cols, n_tables, n_b_cols, name, name_cap, be_rowmajor_out);
}

extern "C" int lsp_cbor_lookup_read_rows(const uint8_t* cbor, size_t len, size_t min_rows, size_t* rows, uint32_t* n_a_cols,
                                         uint32_t* n_tables, uint32_t* n_b_cols, char* name, size_t name_cap, uint8_t** be_rowmajor_out) {
    if (!cbor || !rows || !n_a_cols || !n_tables || !n_b_cols || !be_rowmajor_out) return LSP_ERR_PARAM;
    LookupShape s;
    std::vector<Run> runs;
    if (!lookup_shape(cbor, len, s, &runs) || s.height() == 0) return LSP_ERR_PARAM;
    const size_t h = s.height() > min_rows ? s.height() : min_rows, stride = s.a_rows.size() + s.b_rows.size() * (s.b_rows[0].size() + 1) + 1;
    if (h > SIZE_MAX / 32 / stride) return LSP_ERR_NOMEM;   // a target height whose buffer size would wrap
    uint8_t* out = static_cast<uint8_t*>(host_alloc(h * stride * 32));
    if (!out) return LSP_ERR_NOMEM;
    if (!fill_lookup(cbor, len, s, runs, out, h)) {
        host_release(out);
        return LSP_ERR_PARAM;
    }
    *rows = h;
    *n_a_cols = uint32_t(s.a_rows.size());
    *n_tables = uint32_t(s.b_rows.size());
    *n_b_cols = uint32_t(s.b_rows[0].size());
    copy_name(s.name, name, name_cap);
    *be_rowmajor_out = out;
    return LSP_OK;
}

extern "C" int lsp_cbor_permutation_shape(const uint8_t* cbor, size_t len, size_t* rows, uint32_t* n_cols, char* name, size_t name_cap) {
    if (!cbor || !rows || !n_cols) return LSP_ERR_PARAM;
    Shape s;
    std::vector<Run> runs;
    if (!permutation_shape(cbor, len, s, &runs)) return LSP_ERR_PARAM;
    *rows = s.height();
    *n_cols = uint32_t(s.a_rows.size());
    copy_name(s.name, name, name_cap);
    return LSP_OK;
}

extern "C" int lsp_cbor_permutation_decode(const uint8_t* cbor, size_t len, uint8_t* be_rowmajor, size_t rows, uint32_t n_cols) {
    if (!cbor || !be_rowmajor || rows == 0 || n_cols == 0) return LSP_ERR_PARAM;
    Shape s;
    std::vector<Run> runs;
    if (!permutation_shape(cbor, len, s, &runs)) return LSP_ERR_PARAM;
    if (s.a_rows.size() != n_cols || s.height() > rows) return LSP_ERR_PARAM;
    return fill_permutation(cbor, len, s, runs, be_rowmajor, rows) ? LSP_OK : LSP_ERR_PARAM;
}

extern "C" int lsp_cbor_permutation_read(const uint8_t* cbor, size_t len, size_t* rows, uint32_t* n_cols, char* name, size_t name_cap,
                                         uint8_t** be_rowmajor_out) {
    return lsp_cbor_permutation_read_rows(cbor, len, 0, rows, n_cols, name, name_cap, be_rowmajor_out);
}

extern "C" int lsp_cbor_permutation_read_rows(const uint8_t* cbor, size_t len, size_t min_rows, size_t* rows, uint32_t* n_cols, char* name,
                                              size_t name_cap, uint8_t** be_rowmajor_out) {
    if (!cbor || !rows || !n_cols || !be_rowmajor_out) return LSP_ERR_PARAM;
    Shape s;
    std::vector<Run> runs;
    if (!permutation_shape(cbor, len, s, &runs) || s.height() == 0) return LSP_ERR_PARAM;
    const size_t h = s.height() > min_rows ? s.height() : min_rows, nc = s.a_rows.size();
    if (h > SIZE_MAX / 64 / nc) return LSP_ERR_NOMEM;       // a target height whose buffer size would wrap
    uint8_t* out = static_cast<uint8_t*>(host_alloc(h * 2 * nc * 32));
    if (!out) return LSP_ERR_NOMEM;
    if (!fill_permutation(cbor, len, s, runs, out, h)) {
        host_release(out);
        return LSP_ERR_PARAM;
    }
    *rows = h;
    *n_cols = uint32_t(nc);
    copy_name(s.name, name, name_cap);
    *be_rowmajor_out = out;
    return LSP_OK;
}

extern "C" void lsp_host_free(void* p) { host_release(p); }
