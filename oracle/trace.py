"""Witness generation for the permutation argument -- oracle restatement.

Follows `trace/src/permutation.rs:24-118` (`RawPermutationTrace::get_trace`,
`get_columns`), `trace/src/lib.rs:50-106` (`push_permutation`, `get_trace`)
and the CBOR wire format of `trace/src/permutation.rs:9-22`
(SURVEY.md A.12).  `synthetic_permutation_input` is the seeded workload
generator SURVEY.md 8(d) specifies.
"""
from __future__ import annotations

from .air import AirPermutationConfig
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


def row_major(cols):
    """`RawTrace::get_trace` (`trace/src/lib.rs:94-106`)."""
    h = len(cols[0])
    return [[cols[c][r] for c in range(len(cols))] for r in range(h)]


def build_trace(perm_inputs, alpha: int, delta: int):
    """`RawTrace::push_traces` for permutation traces only (`trace/src/lib.rs:62-92`):
    pads every input to the tallest, concatenates columns, shifts configs."""
    height = max(max(len(col) for col in a + b) for a, b in perm_inputs)
    cols, cfgs = [], []
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
