#!/usr/bin/env python
"""Generates tests/golden/golden_verify_v1.json: rejection reasons of the verifier on damaged copies of the first
proof of golden_v1.json (the one stored word for word).

    python tests/golden/make_golden_verify.py

Each vector is (word index, bit to flip, expected code); the codes are those of include/lsp_b200.h (LSP_VERIFY_*) =
oracle/c/lsp_oracle.c, produced here by the C port and cross-checked against the Python oracle's error names wherever
the damage leaves the sampled query indices equal to the stored ones (its structured proof has no index words).  Like
golden_v1.json this freezes the ORACLE's behaviour (the reference ships no such vectors)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import air as OA  # noqa: E402
from oracle import cport  # noqa: E402
from oracle import field as F  # noqa: E402
from oracle import poseidon2 as OP  # noqa: E402
from oracle import stark as OS  # noqa: E402
from tests.proofs import dict_from_flat  # noqa: E402
from tests.test_golden import GOLD, instance, ix  # noqa: E402

NAMES = {"InvalidProofShape": {1}, "InputError(MerkleRootMismatch)": {2, 3}, "CommitPhaseMmcsError": {4},
         "FinalPolyMismatch": {5}, "InvalidPowWitness": {6}, "OodEvaluationMismatch": {7}}


def main():
    case = GOLD["proofs"][0]
    assert "flat_words_hex" in case
    p = OP.Poseidon2Params.from_seed(0xB200, sbox_d=5)
    cport.set_poseidon2(p)
    cfgs, trace, publics = instance(case)
    fri = OS.FriConfig(**case["fri"])
    words = np.array([ix(v) for v in case["flat_words_hex"]], dtype=np.uint64)
    w, log_n = OA.air_width(cfgs), case["log_n"]
    pub = np.array([F.to_mont_limbs(x) for x in publics], dtype=np.uint64)
    assert cport.verify_limbs(fri, log_n, w, cfgs, pub, words) == 0
    rng = np.random.default_rng(20261018)
    n_elems = words.size // 4
    picks = sorted(set([0, 1, 2, 2 + w, 2 + 2 * w, n_elems - 1] + [int(x) for x in rng.integers(0, n_elems, size=60)]))
    vectors = []
    for e in picks:
        limb = int(rng.integers(0, 3))              # never the top limb: the damaged element stays below r
        bit = int(rng.integers(0, 64))
        bad = words.copy()
        bad[4 * e + limb] ^= np.uint64(1 << bit)
        code = int(cport.verify_limbs(fri, log_n, w, cfgs, pub, bad))
        assert code != 0
        if code != 1:
            # the stored indices matched the sampled ones, so the same damage is expressible for the Python verifier,
            # which must name the same check (code 1 = a stored index differs from the sampled one: not in its proof type)
            d, _ = dict_from_flat(bad, log_n, w, 1, fri)
            try:
                OS.verify(p, fri, cfgs, d, publics)
                name = None
            except OS.VerificationError as err:
                name = str(err)
            assert name is not None and code in NAMES[name], (e, limb, bit, code, name)
        vectors.append([4 * e + limb, bit, code])
    out = Path(__file__).with_name("golden_verify_v1.json")
    out.write_text(json.dumps({"about": "see make_golden_verify.py", "proof": 0, "vectors": vectors}) + "\n")
    print("wrote", out, len(vectors), "vectors; codes seen:", sorted({v[2] for v in vectors}))


if __name__ == "__main__":
    main()
