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
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lsp_b200.h"

namespace {

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;

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

    // Skips one complete item of any type.
    void skip() {
        uint64_t arg;
        bool indef;
        int m = head(arg, indef);
        if (!ok) return;
        switch (m) {
            case 0: case 1: case 7: break;
            case 2: case 3:
                if (indef) {
                    while (ok && !at_break()) skip();
                    if (ok) skip_break();
                } else if (need(arg)) {
                    p += arg;
                }
                break;
            case 4: case 5: {
                uint64_t n = m == 5 ? 2 * arg : arg;
                if (indef) {
                    while (ok && !at_break()) skip();
                    if (ok) skip_break();
                } else {
                    for (uint64_t i = 0; ok && i < n; i++) skip();
                }
                break;
            }
            case 6: skip(); break;
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
    std::string name;
    size_t height() const {
        size_t h = 0;
        for (size_t r : a_rows) h = r > h ? r : h;
        for (size_t r : b_rows) h = r > h ? r : h;
        return h;
    }
};

// Walks `a` or `b`: array(columns) of array(rows) of [u8;32].  With `out` set, column j's row i is
// written at out + ((i * stride_cols) + col0 + j) * 32.
bool walk_columns(Reader& r, std::vector<size_t>* rows_out, uint8_t* out, size_t stride_cols, size_t col0) {
    uint64_t nc;
    bool ic;
    if (r.head(nc, ic) != 4) return r.ok = false;
    size_t j = 0;
    while (r.ok && (ic ? !r.at_break() : j < nc)) {
        uint64_t nr;
        bool ir;
        if (r.head(nr, ir) != 4) return r.ok = false;
        size_t i = 0;
        while (r.ok && (ir ? !r.at_break() : i < nr)) {
            r.elem32(out ? out + ((i * stride_cols) + col0 + j) * 32 : nullptr);
            i++;
        }
        if (r.ok && ir) r.skip_break();
        if (rows_out) rows_out->push_back(i);
        j++;
    }
    if (r.ok && ic) r.skip_break();
    return r.ok;
}

// One pass over the top-level map.  Shape pass: out == nullptr.
bool walk(const uint8_t* cbor, size_t len, Shape* shape, uint8_t* out, size_t stride_cols, size_t n_a) {
    Reader r{cbor, cbor + len};
    uint64_t n;
    bool indef;
    if (r.head(n, indef) != 5) return false;
    bool seen_a = false, seen_b = false;
    size_t k = 0;
    while (r.ok && (indef ? !r.at_break() : k < n)) {
        std::string key;
        if (!r.text(key)) return false;
        if (key == "a") {
            walk_columns(r, shape ? &shape->a_rows : nullptr, out, stride_cols, 0);
            seen_a = true;
        } else if (key == "b") {
            walk_columns(r, shape ? &shape->b_rows : nullptr, out, stride_cols, n_a);
            seen_b = true;
        } else if (key == "name" && shape) {
            r.text(shape->name);
        } else {
            r.skip();
        }
        k++;
    }
    return r.ok && seen_a && seen_b;
}

// ---- RawLookupTrace (trace/src/lookup.rs:10-17) -----------------------------------------------------
//   { a: Vec<Vec<[u8;32]>>, b: Vec<Vec<Vec<[u8;32]>>>, name, a_filter: Vec<[u8;32]>, b_filter: Vec<Vec<[u8;32]>> }
struct LookupShape {
    std::vector<size_t> a_rows;                 // per a column
    std::vector<std::vector<size_t>> b_rows;    // per table, per column
    size_t a_filter_len = 0;
    std::vector<size_t> b_filter_len;           // per table present in the file
    std::string name;
    size_t height() const {                     // get_max_height (:215-228)
        size_t h = 0;
        for (size_t r : a_rows) h = r > h ? r : h;
        for (auto& t : b_rows)
            for (size_t r : t) h = r > h ? r : h;
        return h;
    }
};

// array(rows) of [u8;32]; row i goes to out + (i*stride + col)*32 when out != null and i < max_rows.  Returns the length.
size_t walk_vector(Reader& r, uint8_t* out, size_t stride, size_t col, size_t max_rows) {
    uint64_t nr;
    bool ir;
    if (r.head(nr, ir) != 4) {
        r.ok = false;
        return 0;
    }
    size_t i = 0;
    while (r.ok && (ir ? !r.at_break() : i < nr)) {
        r.elem32(out && i < max_rows ? out + (i * stride + col) * 32 : nullptr);
        i++;
    }
    if (r.ok && ir) r.skip_break();
    return i;
}

// One pass over the top-level map.  Decode pass: out != null, with the shape `sh` of the first pass.
bool walk_lookup(const uint8_t* cbor, size_t len, LookupShape* shape, const LookupShape* sh, uint8_t* out, size_t rows) {
    Reader r{cbor, cbor + len};
    uint64_t n;
    bool indef;
    if (r.head(n, indef) != 5) return false;
    const size_t n_a = sh ? sh->a_rows.size() : 0, n_t = sh ? sh->b_rows.size() : 0, n_b = (sh && n_t) ? sh->b_rows[0].size() : 0;
    const size_t stride = n_a + n_t * n_b + 1 + n_t;
    bool seen_a = false, seen_b = false;
    size_t k = 0;
    while (r.ok && (indef ? !r.at_break() : k < n)) {
        std::string key;
        if (!r.text(key)) return false;
        uint64_t cnt;
        bool ic;
        if (key == "a") {
            if (r.head(cnt, ic) != 4) return false;
            size_t j = 0;
            while (r.ok && (ic ? !r.at_break() : j < cnt)) {
                size_t l = walk_vector(r, out, stride, j, rows);
                if (shape) shape->a_rows.push_back(l);
                j++;
            }
            if (r.ok && ic) r.skip_break();
            seen_a = true;
        } else if (key == "b") {
            if (r.head(cnt, ic) != 4) return false;
            size_t t = 0;
            while (r.ok && (ic ? !r.at_break() : t < cnt)) {
                uint64_t nc;
                bool icc;
                if (r.head(nc, icc) != 4) return false;
                if (shape) shape->b_rows.emplace_back();
                size_t j = 0;
                while (r.ok && (icc ? !r.at_break() : j < nc)) {
                    size_t l = walk_vector(r, out, stride, n_a + t * n_b + j, rows);
                    if (shape) shape->b_rows.back().push_back(l);
                    j++;
                }
                if (r.ok && icc) r.skip_break();
                t++;
            }
            if (r.ok && ic) r.skip_break();
            seen_b = true;
        } else if (key == "a_filter") {
            size_t l = walk_vector(r, out, stride, n_a + n_t * n_b, rows);
            if (shape) shape->a_filter_len = l;
        } else if (key == "b_filter") {
            if (r.head(cnt, ic) != 4) return false;
            size_t t = 0;
            while (r.ok && (ic ? !r.at_break() : t < cnt)) {
                // filters of tables the file does not have are parsed and dropped
                size_t l = walk_vector(r, (out && t < n_t) ? out : nullptr, stride, n_a + n_t * n_b + 1 + t, rows);
                if (shape) shape->b_filter_len.push_back(l);
                t++;
            }
            if (r.ok && ic) r.skip_break();
        } else if (key == "name" && shape) {
            r.text(shape->name);
        } else {
            r.skip();
        }
        k++;
    }
    return r.ok && seen_a && seen_b;
}

bool lookup_shape(const uint8_t* cbor, size_t len, LookupShape& s) {
    if (!walk_lookup(cbor, len, &s, nullptr, nullptr, 0)) return false;
    if (s.a_rows.empty() || s.b_rows.empty() || s.b_rows[0].empty()) return false;
    for (auto& t : s.b_rows)
        if (t.size() != s.b_rows[0].size()) return false;   // AirLookupConfig::width assumes equal table widths (air_lookup.rs:37-39)
    return true;
}

}  // namespace

extern "C" int lsp_cbor_lookup_shape(const uint8_t* cbor, size_t len, size_t* rows, uint32_t* n_a_cols, uint32_t* n_tables,
                                     uint32_t* n_b_cols, char* name, size_t name_cap) {
    if (!cbor || !rows || !n_a_cols || !n_tables || !n_b_cols) return LSP_ERR_PARAM;
    LookupShape s;
    if (!lookup_shape(cbor, len, s)) return LSP_ERR_PARAM;
    *rows = s.height();
    *n_a_cols = uint32_t(s.a_rows.size());
    *n_tables = uint32_t(s.b_rows.size());
    *n_b_cols = uint32_t(s.b_rows[0].size());
    if (name && name_cap) {
        size_t n = s.name.size() < name_cap - 1 ? s.name.size() : name_cap - 1;
        memcpy(name, s.name.data(), n);
        name[n] = 0;
    }
    return LSP_OK;
}

extern "C" int lsp_cbor_lookup_decode(const uint8_t* cbor, size_t len, uint8_t* be_rowmajor, size_t rows, uint32_t n_a_cols,
                                      uint32_t n_tables, uint32_t n_b_cols) {
    if (!cbor || !be_rowmajor || rows == 0) return LSP_ERR_PARAM;
    LookupShape s;
    if (!lookup_shape(cbor, len, s)) return LSP_ERR_PARAM;
    if (s.height() != rows || s.a_rows.size() != n_a_cols || s.b_rows.size() != n_tables || s.b_rows[0].size() != n_b_cols) return LSP_ERR_PARAM;
    const size_t stride = size_t(n_a_cols) + size_t(n_tables) * n_b_cols + 1 + n_tables;
    memset(be_rowmajor, 0, rows * stride * 32);          // `resize` pads columns and filters with zeros (:230-246)
    if (!walk_lookup(cbor, len, nullptr, &s, be_rowmajor, rows)) return LSP_ERR_PARAM;
    // `read_file` (:25-41): missing filter entries default to ONE up to the length of the first column they guard
    auto fill_ones = [&](size_t col, size_t from, size_t to) {
        for (size_t i = from; i < to && i < rows; i++) be_rowmajor[(i * stride + col) * 32 + 31] = 1;
    };
    fill_ones(size_t(n_a_cols) + size_t(n_tables) * n_b_cols, s.a_filter_len, s.a_rows[0]);
    for (uint32_t t = 0; t < n_tables; t++) {
        size_t have = t < s.b_filter_len.size() ? s.b_filter_len[t] : 0;
        fill_ones(size_t(n_a_cols) + size_t(n_tables) * n_b_cols + 1 + t, have, s.b_rows[t][0]);
    }
    return LSP_OK;
}

extern "C" int lsp_cbor_permutation_shape(const uint8_t* cbor, size_t len, size_t* rows, uint32_t* n_cols, char* name, size_t name_cap) {
    if (!cbor || !rows || !n_cols) return LSP_ERR_PARAM;
    Shape s;
    if (!walk(cbor, len, &s, nullptr, 0, 0)) return LSP_ERR_PARAM;
    if (s.a_rows.empty() || s.a_rows.size() != s.b_rows.size()) return LSP_ERR_PARAM;  // air/src/lib.rs zips a and b ids
    *rows = s.height();
    *n_cols = uint32_t(s.a_rows.size());
    if (name && name_cap) {
        size_t n = s.name.size() < name_cap - 1 ? s.name.size() : name_cap - 1;
        memcpy(name, s.name.data(), n);
        name[n] = 0;
    }
    return LSP_OK;
}

extern "C" int lsp_cbor_permutation_decode(const uint8_t* cbor, size_t len, uint8_t* be_rowmajor, size_t rows, uint32_t n_cols) {
    if (!cbor || !be_rowmajor || rows == 0 || n_cols == 0) return LSP_ERR_PARAM;
    size_t r0 = 0;
    uint32_t c0 = 0;
    int rc = lsp_cbor_permutation_shape(cbor, len, &r0, &c0, nullptr, 0);
    if (rc != LSP_OK) return rc;
    if (r0 != rows || c0 != n_cols) return LSP_ERR_PARAM;
    memset(be_rowmajor, 0, rows * size_t(2) * n_cols * 32);  // short columns are zero-padded (`resize`, permutation.rs:134-142)
    return walk(cbor, len, nullptr, be_rowmajor, size_t(2) * n_cols, n_cols) ? LSP_OK : LSP_ERR_PARAM;
}
