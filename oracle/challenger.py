"""`HashChallenger<Val, Hash, 1>` -- oracle restatement.

Reference anchors: `Challenger = HashChallenger<Val,Hash,1>`
(`bin/src/config.rs:23`), constructed with an empty initial state at
`bin/src/main.rs:78,88`.  Algorithm: published Plonky3 `p3-challenger`
(SURVEY.md A.6).  `sample_bits` / `grind` for this 253-bit field are
fork-only; restated as "low bits of the canonical integer" and "smallest
witness" (deterministic where the reference's rayon `find_any` is not).
"""
from __future__ import annotations

from .poseidon2 import Poseidon2Params, hash_iter


class HashChallenger:
    def __init__(self, p: Poseidon2Params, initial_state=()):
        self.p = p
        self.input_buffer = list(initial_state)
        self.output_buffer = []

    def clone(self) -> "HashChallenger":
        c = HashChallenger(self.p, self.input_buffer)
        c.output_buffer = list(self.output_buffer)
        return c

    def _flush(self):
        out = hash_iter(self.p, self.input_buffer)
        self.output_buffer = [out]
        self.input_buffer = [out]  # chaining value

    def observe(self, x: int):
        self.output_buffer = []
        self.input_buffer.append(x)

    def observe_slice(self, xs):
        for x in xs:
            self.observe(x)

    def sample(self) -> int:
        if not self.output_buffer:
            self._flush()
        return self.output_buffer.pop()

    def sample_bits(self, bits: int) -> int:
        return self.sample() & ((1 << bits) - 1)

    def check_witness(self, bits: int, witness: int) -> bool:
        self.observe(witness)
        return self.sample_bits(bits) == 0

    def grind(self, bits: int) -> int:
        w = 0
        while not self.clone().check_witness(bits, w):
            w += 1
        assert self.check_witness(bits, w)
        return w
