"""MerkleTreeMmcs<Val,Val,Hash,Compress,1> -- oracle restatement.

Reference anchors: `ValMmcs` / `ChallengeMmcs` (`bin/src/config.rs:19-20`,
built at `bin/src/main.rs:56-57`).  Algorithm: published Plonky3
`p3-merkle-tree`.  Every commit on the prover path holds
matrices of one common height, so the mixed-height injection rule is not
restated; mixed heights raise.
"""
from __future__ import annotations

from .field import log2_strict
from .poseidon2 import Poseidon2Params, compress, hash_iter


class MerkleTree:
    def __init__(self, p: Poseidon2Params, mats):
        """mats: list of row-major matrices, each a list of rows (lists of ints)."""
        assert len(mats) > 0
        h = len(mats[0])
        if any(len(m) != h for m in mats):
            raise ValueError("mixed-height commit is not on the reference's path")
        log2_strict(h)
        self.p = p
        self.mats = mats
        self.height = h
        layer = []
        for i in range(h):
            row = []
            for m in mats:
                row += m[i]
            layer.append(hash_iter(p, row))
        self.layers = [layer]
        while len(layer) > 1:
            layer = [compress(p, layer[2 * i], layer[2 * i + 1]) for i in range(len(layer) // 2)]
            self.layers.append(layer)

    @property
    def root(self) -> int:
        return self.layers[-1][0]

    def open_batch(self, index: int):
        rows = [list(m[index]) for m in self.mats]
        proof = [self.layers[k][(index >> k) ^ 1] for k in range(len(self.layers) - 1)]
        return rows, proof


def verify_batch(p: Poseidon2Params, root: int, height: int, index: int, rows, proof) -> bool:
    if len(proof) != log2_strict(height):
        return False
    flat = []
    for r in rows:
        flat += r
    node = hash_iter(p, flat)
    for k, sib in enumerate(proof):
        if (index >> k) & 1:
            node = compress(p, sib, node)
        else:
            node = compress(p, node, sib)
    return node == root
