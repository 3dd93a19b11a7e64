"""Poseidon2 (width 3) + sponge + 2-to-1 compression -- oracle restatement.

Reference anchors: `Perm = Poseidon2Bls12337<3>` (`bin/src/config.rs:11`),
`Perm::new_from_rng(8, 22, &mut rng)` (`bin/src/main.rs:49`),
`Hash = PaddingFreeSponge<Perm,3,2,1>` (`bin/src/config.rs:12`),
`Compress = CompressionFunctionFromHasher<Hash,2,1>` (`bin/src/config.rs:17`).
Algorithms: published Plonky3 `p3-poseidon2` / `p3-symmetric`
(SURVEY.md A.4, A.5).  The S-box degree and the internal diagonal are
parameters because the fork-only crate `p3-bls12-377-fr` is not available.
"""
from __future__ import annotations

from dataclasses import dataclass, field as dc_field

from .field import R_MOD, SplitMix64

WIDTH = 3
RATE = 2


@dataclass
class Poseidon2Params:
    sbox_d: int = 5
    rounds_f: int = 8
    rounds_p: int = 22
    ext_initial: list = dc_field(default_factory=list)   # rounds_f/2 x 3
    ext_terminal: list = dc_field(default_factory=list)  # rounds_f/2 x 3
    internal: list = dc_field(default_factory=list)      # rounds_p
    internal_diag_m1: tuple = (1, 1, 2)                  # bn254-style width-3 diagonal

    @staticmethod
    def from_seed(seed: int, sbox_d: int = 5, rounds_f: int = 8, rounds_p: int = 22) -> "Poseidon2Params":
        """Draw order of `new_from_rng`: initial external
        constants, terminal external constants, then internal constants."""
        rng = SplitMix64(seed)
        half = rounds_f // 2
        ini = [[rng.next_fr() for _ in range(WIDTH)] for _ in range(half)]
        ter = [[rng.next_fr() for _ in range(WIDTH)] for _ in range(half)]
        internal = [rng.next_fr() for _ in range(rounds_p)]
        return Poseidon2Params(sbox_d, rounds_f, rounds_p, ini, ter, internal)

    def flat_constants(self) -> list:
        """[ext_initial..., ext_terminal..., internal...] -- the order the C ABI takes."""
        out = []
        for rc in self.ext_initial:
            out += rc
        for rc in self.ext_terminal:
            out += rc
        out += self.internal
        return out


def _ext_linear(s):
    # M_E = circ(2,1,1): s_i += sum(s)  (mds_light_permutation, WIDTH=3)
    t = (s[0] + s[1] + s[2]) % R_MOD
    return [(s[0] + t) % R_MOD, (s[1] + t) % R_MOD, (s[2] + t) % R_MOD]


def permute(p: Poseidon2Params, state):
    s = [x % R_MOD for x in state]
    d = p.sbox_d
    s = _ext_linear(s)
    for rc in p.ext_initial:
        s = [pow((s[i] + rc[i]) % R_MOD, d, R_MOD) for i in range(WIDTH)]
        s = _ext_linear(s)
    dg = p.internal_diag_m1
    for rc in p.internal:
        s[0] = pow((s[0] + rc) % R_MOD, d, R_MOD)
        t = (s[0] + s[1] + s[2]) % R_MOD
        s = [(s[i] * dg[i] + t) % R_MOD for i in range(WIDTH)]
    for rc in p.ext_terminal:
        s = [pow((s[i] + rc[i]) % R_MOD, d, R_MOD) for i in range(WIDTH)]
        s = _ext_linear(s)
    return s


def hash_iter(p: Poseidon2Params, xs) -> int:
    """PaddingFreeSponge<Perm,3,2,1>::hash_iter: overwrite-mode, rate 2, out = state[0]."""
    state = [0, 0, 0]
    it = iter(xs)
    while True:
        for i in range(RATE):
            try:
                state[i] = next(it) % R_MOD
            except StopIteration:
                if i != 0:
                    state = permute(p, state)
                return state[0]
        state = permute(p, state)


def compress(p: Poseidon2Params, l: int, r: int) -> int:
    """CompressionFunctionFromHasher<Hash,2,1>::compress([l],[r])."""
    return hash_iter(p, [l, r])


def num_perms_for_row(width: int) -> int:
    return (width + RATE - 1) // RATE
