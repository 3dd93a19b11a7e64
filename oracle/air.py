"""LineaAIR permutation (grand-product) and lookup (LogUp) constraints -- oracle restatement.

Follows `air/src/lib.rs:116-167` (`eval_permutation`), `air/src/lib.rs:57-114` (`eval_lookup`), the
column-id configs `air/src/air_permutation.rs:2-23` / `air/src/air_lookup.rs:2-39` and the dispatch
loop `air/src/lib.rs:47-54`.
`eval` is written once against a tiny expression algebra so that the same
code serves the prover folder, the verifier folder, the debug
`check_constraints` and the symbolic degree inference.
"""
from __future__ import annotations

from dataclasses import dataclass

from .field import R_MOD, log2_ceil


@dataclass
class AirPermutationConfig:
    """`air/src/air_permutation.rs:2-7`."""
    a_columns_ids: list
    b_columns_ids: list
    b_inverse_id: int
    check_id: int

    def shift(self, s: int):  # `air/src/air_permutation.rs:10-19`
        self.a_columns_ids = [i + s for i in self.a_columns_ids]
        self.b_columns_ids = [i + s for i in self.b_columns_ids]
        self.b_inverse_id += s
        self.check_id += s

    def width(self) -> int:  # `air/src/air_permutation.rs:21-23`
        return len(self.a_columns_ids) + len(self.b_columns_ids) + 2

    @staticmethod
    def standard(c: int, offset: int = 0) -> "AirPermutationConfig":
        """Ids emitted by `RawPermutationTrace::get_trace` (`trace/src/permutation.rs:84-92`)."""
        cfg = AirPermutationConfig(list(range(c)), list(range(c, 2 * c)), 2 * c, 2 * c + 1)
        cfg.shift(offset)
        return cfg


@dataclass
class AirLookupConfig:
    """`air/src/air_lookup.rs:2-11`."""
    a_columns_ids: list
    b_columns_ids: list          # list (tables) of lists (columns)
    a_filter_id: int
    b_filter_id: list
    a_inverses_id: int
    b_inverses_id: list
    occurrences_id: list
    check_id: int

    def shift(self, s: int):  # `air/src/air_lookup.rs:14-35`
        self.a_columns_ids = [i + s for i in self.a_columns_ids]
        self.b_columns_ids = [[i + s for i in t] for t in self.b_columns_ids]
        self.a_filter_id += s
        self.b_filter_id = [i + s for i in self.b_filter_id]
        self.a_inverses_id += s
        self.b_inverses_id = [i + s for i in self.b_inverses_id]
        self.occurrences_id = [i + s for i in self.occurrences_id]
        self.check_id += s

    def width(self) -> int:  # `air/src/air_lookup.rs:37-39`
        return len(self.a_columns_ids) + len(self.b_columns_ids) * (len(self.b_columns_ids[0]) + 3) + 3

    @staticmethod
    def standard(n_a: int, n_tables: int, n_b: int, offset: int = 0) -> "AirLookupConfig":
        """Ids emitted by `RawLookupTrace::get_air_lookup_config` (`trace/src/lookup.rs:178-214`)."""
        a_ids = list(range(n_a))
        b_ids = [[n_a + t * n_b + k for k in range(n_b)] for t in range(n_tables)]
        a_filter = b_ids[-1][-1] + 1
        b_filter = [a_filter + 1 + t for t in range(n_tables)]
        a_inv = b_filter[-1] + 1
        b_inv = [a_inv + 1 + t for t in range(n_tables)]
        occ = [b_inv[-1] + 1 + t for t in range(n_tables)]
        cfg = AirLookupConfig(a_ids, b_ids, a_filter, b_filter, a_inv, b_inv, occ, occ[-1] + 1)
        cfg.shift(offset)
        return cfg


class Fe:
    """Field element with operators (prover / verifier folders)."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v % R_MOD

    def __add__(self, o):
        return Fe(self.v + o.v)

    def __sub__(self, o):
        return Fe(self.v - o.v)

    def __mul__(self, o):
        return Fe(self.v * o.v)


class Deg:
    """`SymbolicExpression::degree_multiple`: trace var 1, public/constant 0,
    is_first_row/is_last_row 1, is_transition 0; mul adds, add/sub max."""
    __slots__ = ("d",)

    def __init__(self, d):
        self.d = d

    def __add__(self, o):
        return Deg(max(self.d, o.d))

    __sub__ = __add__

    def __mul__(self, o):
        return Deg(self.d + o.d)


def eval_permutation(cfg: AirPermutationConfig, local, nxt, alpha, delta, zero, one,
                     is_first, is_last, is_transition):
    """Returns the constraints in emission order (`air/src/lib.rs:116-167`);
    `when_*` multiplies by the selector, `assert_eq(x,y)` asserts x - y."""
    a_local = zero
    for i in cfg.a_columns_ids:            # :129-132
        a_local = a_local * alpha + local[i]
    b_local = zero
    for i in cfg.b_columns_ids:            # :134-137
        b_local = b_local * alpha + local[i]
    a_ch = a_local + delta                 # :139
    b_ch = b_local + delta                 # :140
    out = []
    out.append(b_ch * local[cfg.b_inverse_id] - one)                                   # :143
    out.append(is_first * (local[cfg.check_id] - a_ch * local[cfg.b_inverse_id]))      # :146-148
    a_next = zero
    for i in cfg.a_columns_ids:            # :150-153
        a_next = a_next * alpha + nxt[i]
    a_next_ch = a_next + delta             # :155
    out.append(is_transition * (nxt[cfg.check_id]
                                - local[cfg.check_id] * a_next_ch * nxt[cfg.b_inverse_id]))  # :158-161
    out.append(is_last * (local[cfg.check_id] - one))                                  # :164-166
    return out


def eval_lookup(cfg: AirLookupConfig, local, nxt, alpha, delta, zero, one, is_first, is_last, is_transition):
    """Constraints of `eval_lookup` in emission order (`air/src/lib.rs:57-114`)."""
    a_local = zero
    for i in cfg.a_columns_ids:                                    # :65-68
        a_local = a_local * alpha + local[i]
    a_ch = a_local + delta                                         # :70
    out = [a_ch * local[cfg.a_inverses_id] - one]                  # :73
    local_check = local[cfg.a_filter_id] * local[cfg.a_inverses_id]   # :75
    next_check = nxt[cfg.a_filter_id] * nxt[cfg.a_inverses_id]        # :76
    for t, b_ids in enumerate(cfg.b_columns_ids):                  # :78
        b_local = zero
        for i in b_ids:                                            # :79-82
            b_local = b_local * alpha + local[i]
        b_ch = b_local + delta                                     # :84
        out.append(b_ch * local[cfg.b_inverses_id[t]] - one)       # :85-88
        local_check = local_check - local[cfg.b_filter_id[t]] * local[cfg.occurrences_id[t]] * local[cfg.b_inverses_id[t]]  # :90-92
        next_check = next_check - nxt[cfg.b_filter_id[t]] * nxt[cfg.occurrences_id[t]] * nxt[cfg.b_inverses_id[t]]         # :94-96
    out.append(is_first * (local[cfg.check_id] - local_check))                         # :100-102
    out.append(is_transition * ((nxt[cfg.check_id] - local[cfg.check_id]) - next_check))  # :105-107
    out.append(is_last * (local[cfg.check_id] - zero))                                 # :110-112
    return out


def eval_air(cfgs, local, nxt, publics, zero, one, is_first, is_last, is_transition):
    """`LineaAIR::eval` (`air/src/lib.rs:47-54`); publics = [alpha, delta] (`bin/src/main.rs:85`)."""
    out = []
    for c in cfgs:
        f = eval_lookup if isinstance(c, AirLookupConfig) else eval_permutation
        out += f(c, local, nxt, publics[0], publics[1], zero, one, is_first, is_last, is_transition)
    return out


def air_width(cfgs) -> int:  # `air/src/lib.rs:33-38`
    return sum(c.width() for c in cfgs)


def log_quotient_degree(cfgs) -> int:
    """`get_log_quotient_degree` of p3-uni-stark."""
    w = air_width(cfgs)
    tv = [Deg(1)] * w
    cs = eval_air(cfgs, tv, tv, [Deg(0), Deg(0)], Deg(0), Deg(0), Deg(1), Deg(1), Deg(0))
    d = max([c.d for c in cs] + [2])
    return log2_ceil(d - 1)


def num_constraints(cfgs) -> int:
    return sum(4 + len(c.b_columns_ids) if isinstance(c, AirLookupConfig) else 4 for c in cfgs)


def fold_constraints(cfgs, local, nxt, publics, is_first, is_last, is_transition, alpha_stark) -> int:
    """ProverConstraintFolder / VerifierConstraintFolder: acc = acc*alpha + C_k."""
    cs = eval_air(cfgs, [Fe(x) for x in local], [Fe(x) for x in nxt], [Fe(x) for x in publics],
                  Fe(0), Fe(1), Fe(is_first), Fe(is_last), Fe(is_transition))
    acc = 0
    for c in cs:
        acc = (acc * alpha_stark + c.v) % R_MOD
    return acc


def check_constraints(cfgs, trace, publics) -> bool:
    """Debug-build `check_constraints`."""
    n = len(trace)
    for i in range(n):
        cs = eval_air(cfgs, [Fe(x) for x in trace[i]], [Fe(x) for x in trace[(i + 1) % n]],
                      [Fe(x) for x in publics], Fe(0), Fe(1),
                      Fe(1 if i == 0 else 0), Fe(1 if i == n - 1 else 0), Fe(0 if i == n - 1 else 1))
        if any(c.v != 0 for c in cs):
            return False
    return True
