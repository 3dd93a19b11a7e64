"""Witness generation for the permutation argument -- oracle restatement.

Follows `trace/src/permutation.rs:24-118` (`RawPermutationTrace::get_trace`,
`get_columns`), `trace/src/lib.rs:50-106` (`push_permutation`, `get_trace`)
and the CBOR wire format of `trace/src/permutation.rs:9-22`
(SURVEY.md A.12).  `synthetic_permutation_input` is the seeded workload
generator SURVEY.md 8(d) specifies.
"""
from __future__ import annotations

from .air import AirLookupConfig, AirPermutationConfig
from .field import R_MOD, SplitMix64, from_be_bytes_mod_order, inv


def synthetic_permutation_input(seed: int, c: int, n: int, small: bool = False):
    """Returns (a_cols, b_cols): c columns of n elements each; b = the rows of
    `a` shuffled as whole rows by a seeded Fisher-Yates permutation."""
    rng = SplitMix64(seed)
    if small:
        a = [[rng.next_u64() & 0xFFFFFFFF for _ in range(n)] for _ in range(c)]
    else:
        a = [[rng.next_fr() for _ in range(n)] for _ in range(c)]
    perm = list(range(n))
    for i in range(n - 1, 0, -1):
        j = rng.next_below(i + 1)
        perm[i], perm[j] = perm[j], perm[i]
    b = [[a[j][perm[i]] for i in range(n)] for j in range(c)]
    return a, b


def permutation_columns(a, b, alpha: int, delta: int):
    """`RawPermutationTrace::get_trace` (`trace/src/permutation.rs:24-93`):
    columns a.., b.., 1/(b_comb+delta), running product."""
    sz = len(a[0])
    width = len(a)
    prev = 1
    b_inv_col, check_col = [], []
    for i in range(sz):
        a_comb = 0
        for col in a:                      # :56-61
            a_comb = (a_comb * alpha + col[i]) % R_MOD
        b_comb = 0
        for col in b:                      # :63-68
            b_comb = (b_comb * alpha + col[i]) % R_MOD
        bi = inv((b_comb + delta) % R_MOD)  # :70
        b_inv_col.append(bi)
        prev = prev * ((a_comb + delta) % R_MOD) % R_MOD * bi % R_MOD  # :72
        check_col.append(prev)
    if check_col[-1] != 1:                 # :76-79
        raise AssertionError("failed to check constrain: check column should be 1 on the last row")
    cols = [list(x) for x in a] + [list(x) for x in b] + [b_inv_col, check_col]
    return AirPermutationConfig.standard(width), cols


def synthetic_lookup_input(seed: int, n_cols: int, n_tables: int, n: int, table_rows: int = 0, disabled_every: int = 0):
    """A valid LogUp instance: (a, b, a_filter, b_filter) with a = n_cols columns of n looked-up rows,
    b = n_tables tables of n_cols columns of n rows.  The `table_rows` distinct table rows (default
    n//2, at least 1) are dealt round-robin over the tables, each table is padded with copies of its
    own rows, and every A row is drawn from the table rows.  `disabled_every` > 0 switches the
    filter off on every such A row and replaces that row by junk, which the argument must ignore."""
    rng = SplitMix64(seed)
    m = max(1, table_rows or n // 2)
    m = min(m, n * n_tables)
    rows = [[rng.next_fr() for _ in range(n_cols)] for _ in range(m)]
    per_table = [[rows[i] for i in range(t, m, n_tables)] for t in range(n_tables)]
    b = []
    for t in range(n_tables):
        src = per_table[t] or [rows[0]]
        filled = [src[i % len(src)] for i in range(n)]
        b.append([[filled[i][k] for i in range(n)] for k in range(n_cols)])
    picks = [rows[rng.next_below(m)] for _ in range(n)]
    a_filter = [1] * n
    if disabled_every:
        for i in range(0, n, disabled_every):
            a_filter[i] = 0
            picks[i] = [rng.next_fr() for _ in range(n_cols)]
    a = [[picks[i][k] for i in range(n)] for k in range(n_cols)]
    b_filter = [[1] * n for _ in range(n_tables)]
    return a, b, a_filter, b_filter


def lookup_columns(a, b, a_filter, b_filter, alpha: int, delta: int):
    """`RawLookupTrace::get_trace` (`trace/src/lookup.rs:46-176`): columns a.., b (table by table).., a_filter,
    b_filters, 1/(a_comb+delta), per-table 1/(b_comb+delta), per-table multiplicities, running log-derivative sum."""
    sz = len(a[0])
    comb = lambda cols, i: _horner(cols, i, alpha)
    occurrences = {}
    for i in range(sz):                                   # :83-104
        if a_filter[i] == 0:
            continue
        k = comb(a, i)
        occurrences[k] = occurrences.get(k, 0) + 1
    a_inv_col, b_inv_tab, mult_tab, prefix = [], [[] for _ in b], [[] for _ in b], []
    total = 0
    for i in range(sz):                                   # :120-163
        ai = inv((comb(a, i) + delta) % R_MOD)
        a_inv_col.append(ai)
        if a_filter[i] != 0:
            total = (total + ai) % R_MOD
        for t, table in enumerate(b):
            bc = comb(table, i)
            bi = inv((bc + delta) % R_MOD)
            b_inv_tab[t].append(bi)
            occ = 0
            if bc in occurrences and b_filter[t][i] != 0:
                occ = occurrences.pop(bc)
                total = (total - bi * occ) % R_MOD
            mult_tab[t].append(occ)
        prefix.append(total)
    if prefix[-1] != 0:                                   # :165-168
        raise AssertionError("failed to check constrain: check column should be 0 on the last row")
    cols = [list(x) for x in a]
    for table in b:
        cols += [list(x) for x in table]
    cols += [list(a_filter)] + [list(f) for f in b_filter] + [a_inv_col] + b_inv_tab + mult_tab + [prefix]
    return AirLookupConfig.standard(len(a), len(b), len(b[0])), cols


def _horner(cols, i, alpha):
    acc = 0
    for col in cols:
        acc = (acc * alpha + col[i]) % R_MOD
    return acc


def row_major(cols):
    """`RawTrace::get_trace` (`trace/src/lib.rs:94-106`)."""
    h = len(cols[0])
    return [[cols[c][r] for c in range(len(cols))] for r in range(h)]


def build_trace(perm_inputs, alpha: int, delta: int, lookup_inputs=()):
    """`RawTrace::push_traces` (`trace/src/lib.rs:62-92`): the height is the tallest column of ANY input (:67-79);
    lookup traces first, then permutation traces, each resized to that height with zero rows before its witness is
    built (`push_lookup` :37-41 with `resize` `trace/src/lookup.rs:230-246`, filters included, so padded rows are
    disabled; `push_permutation` :50-53 with `trace/src/permutation.rs:134-142`); columns concatenated, configs shifted."""
    heights = [max(len(col) for col in a + b) for a, b in perm_inputs]
    heights += [max([len(col) for col in l[0]] + [len(col) for t in l[1] for col in t]) for l in lookup_inputs]
    height = max(heights)
    pad = lambda col: list(col) + [0] * (height - len(col))
    cols, cfgs = [], []
    for a, b, af, bf in lookup_inputs:
        a, af = [pad(col) for col in a], pad(af)
        b, bf = [[pad(col) for col in t] for t in b], [pad(f) for f in bf]
        cfg, lc = lookup_columns(a, b, af, bf, alpha, delta)
        cfg.shift(len(cols))
        cols += lc
        cfgs.append(cfg)
    for a, b in perm_inputs:
        a = [col + [0] * (height - len(col)) for col in a]   # `resize`, permutation.rs:134-142
        b = [col + [0] * (height - len(col)) for col in b]
        cfg, pc = permutation_columns(a, b, alpha, delta)
        cfg.shift(len(cols))
        cols += pc
        cfgs.append(cfg)
    return cfgs, row_major(cols)


def encode_raw_permutation_trace(a, b, name: str) -> bytes:
    """CBOR of `RawPermutationTrace` (serde: `[u8;32]` is a 32-tuple of small ints)."""
    import cbor2
    enc = lambda cols: [[list(int(x).to_bytes(32, "big")) for x in col] for col in cols]
    return cbor2.dumps({"a": enc(a), "b": enc(b), "name": name})


def decode_raw_permutation_trace(blob: bytes):
    """`read_file` + `get_columns` (`trace/src/permutation.rs:17-22,95-118`)."""
    import cbor2
    obj = cbor2.loads(blob)
    dec = lambda cols: [[from_be_bytes_mod_order(bytes(x)) for x in col] for col in cols]
    return dec(obj["a"]), dec(obj["b"]), obj["name"]


def encode_raw_lookup_trace(a, b, a_filter, b_filter, name: str) -> bytes:
    """CBOR of `RawLookupTrace` (`trace/src/lookup.rs:10-17`); filters may be shorter than the columns or absent
    (`[]`): `read_file` fills them in."""
    import cbor2
    e = lambda x: list(int(x).to_bytes(32, "big"))
    return cbor2.dumps({"a": [[e(x) for x in col] for col in a], "b": [[[e(x) for x in col] for col in t] for t in b], "name": name,
                        "a_filter": [e(x) for x in a_filter], "b_filter": [[e(x) for x in f] for f in b_filter]})


def decode_raw_lookup_trace(blob: bytes, height: int | None = None):
    """`read_file` (`trace/src/lookup.rs:20-44`: default filters), `resize` to `height` (default: the trace's own
    max height, :215-246) and `get_columns` (:248-301).  Returns (a, b, a_filter, b_filter, name), canonical ints."""
    import cbor2
    obj = cbor2.loads(blob)
    d = lambda x: from_be_bytes_mod_order(bytes(x))
    a = [[d(x) for x in col] for col in obj["a"]]
    b = [[[d(x) for x in col] for col in t] for t in obj["b"]]
    af = [d(x) for x in obj.get("a_filter", [])]
    bf = [[d(x) for x in f] for f in obj.get("b_filter", [])]
    af += [1] * (len(a[0]) - len(af))                      # :29-31
    bf += [[] for _ in range(len(b) - len(bf))]            # :33-35
    for t in range(len(b)):                                # :37-41
        bf[t] += [1] * (len(b[t][0]) - len(bf[t]))
    h = height if height is not None else max([len(c) for c in a] + [len(c) for t in b for c in t])
    pad = lambda v: (v + [0] * (h - len(v)))[:h]
    a = [pad(c) for c in a]
    b = [[pad(c) for c in t] for t in b]
    return a, b, pad(af), [pad(f) for f in bf[:len(b)]], obj["name"]
