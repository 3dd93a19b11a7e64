#!/usr/bin/env python
"""Generates tests/golden/golden_v1.json from the Python big-integer oracle.

    python tests/golden/make_golden.py          # rewrites the fixture (run in the build container)

The reference ships no known-answer tests and its arithmetic (the Plonky3 fork, ark-ff) cannot be
built or imported here, so these vectors do NOT come from the reference: they
freeze the oracle's output so that the C port (oracle/c), the CUDA library and any later edit of the
oracle itself are all held to the same numbers.  Values are canonical integers in hex.
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import air as OA  # noqa: E402
from oracle import dft as OD  # noqa: E402
from oracle import field as F  # noqa: E402
from oracle import merkle as OM  # noqa: E402
from oracle import poseidon2 as OP  # noqa: E402
from oracle import stark as OS  # noqa: E402
from oracle import trace as OT  # noqa: E402
from tests.proofs import flat_from_dict  # noqa: E402


def hx(x):
    return format(x, "x")


def proof_case(log_n, c, fri_kw, seed, full):
    p = OP.Poseidon2Params.from_seed(0xB200, sbox_d=5)
    rng = F.SplitMix64(seed)
    alpha, delta = rng.next_fr(), rng.next_fr()
    cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(seed, c, 1 << log_n)], alpha, delta)
    fri = OS.FriConfig(**fri_kw)
    dbg = {}
    proof = OS.prove(p, fri, cfgs, trace, [alpha, delta], dbg)
    OS.verify(p, fri, cfgs, proof, [alpha, delta])
    words = flat_from_dict(proof, dbg["query_indices"])
    out = {"log_n": log_n, "cols": c, "fri": fri_kw, "seed": seed, "alpha": hx(alpha), "delta": hx(delta),
           "trace_commit": hx(proof["commitments"]["trace"]), "quotient_commit": hx(proof["commitments"]["quotient_chunks"]),
           "query_indices": list(dbg["query_indices"]), "n_words": int(words.size),
           "sha256_flat_le_u64": hashlib.sha256(words.astype("<u8").tobytes()).hexdigest()}
    if full:
        out["flat_words_hex"] = [hx(int(w)) for w in words]
    return out


def main():
    rng = F.SplitMix64(0x601D)
    g = {"about": "frozen output of oracle/*.py (see make_golden.py); canonical integers, hex",
         "modulus": hx(F.R_MOD), "generator": hx(F.GENERATOR), "two_adic_root_2_47": hx(F.two_adic_generator(47)),
         "mont_r": hx(F.MONT_R), "mont_r2": hx(F.MONT_R2)}
    edge = [0, 1, 2, F.R_MOD - 1, F.R_MOD - 2, F.MONT_R, (1 << 252), 0xFFFFFFFF, (1 << 64) - 1, F.R_MOD // 2]
    xs = edge + [rng.next_fr() for _ in range(22)]
    ys = list(reversed(edge)) + [rng.next_fr() for _ in range(22)]
    g["field"] = [{"a": hx(a), "b": hx(b), "add": hx(F.add(a, b)), "sub": hx(F.sub(a, b)), "mul": hx(F.mul(a, b)),
                   "inv_a": hx(F.inv(a)) if a else "0", "halve_a": hx(F.halve(a))} for a, b in zip(xs, ys)]
    g["poseidon2"] = []
    for d in (3, 5, 7, 11, 17):
        p = OP.Poseidon2Params.from_seed(0xB200, sbox_d=d)
        states = [[0, 0, 0], [1, 2, 3], [F.R_MOD - 1] * 3] + [[rng.next_fr() for _ in range(3)] for _ in range(3)]
        g["poseidon2"].append({"seed": 0xB200, "sbox_d": d, "rounds_f": 8, "rounds_p": 22,
                               "cases": [{"in": [hx(v) for v in s], "out": [hx(v) for v in OP.permute(p, s)]} for s in states]})
    p5 = OP.Poseidon2Params.from_seed(0xB200, sbox_d=5)
    g["sponge"] = []
    for w in range(0, 8):
        row = [rng.next_fr() for _ in range(w)]
        g["sponge"].append({"row": [hx(v) for v in row], "digest": hx(OP.hash_iter(p5, row))})
    mat = [[rng.next_fr() for _ in range(3)] for _ in range(8)]
    lde = OD.coset_lde_batch(mat, 2, F.GENERATOR)   # already the bit-reversed storage order the PCS commits to
    g["lde"] = {"in": [[hx(v) for v in r] for r in mat], "added_bits": 2, "shift": hx(F.GENERATOR),
                "out_bitrev_storage": [[hx(v) for v in r] for r in lde]}
    tree = OM.MerkleTree(p5, [lde])
    g["merkle"] = {"leaves": "lde.out_bitrev_storage", "root": hx(tree.root), "layers": [[hx(v) for v in l] for l in tree.layers],
                   "open_5": {"siblings": [hx(v) for v in tree.open_batch(5)[1]]}}
    g["proofs"] = [
        proof_case(3, 1, dict(log_blowup=1, log_final_poly_len=0, num_queries=3, proof_of_work_bits=0), 21, True),
        proof_case(4, 3, dict(log_blowup=3, log_final_poly_len=0, num_queries=33, proof_of_work_bits=0), 22, False),
        proof_case(6, 2, dict(log_blowup=2, log_final_poly_len=1, num_queries=5, proof_of_work_bits=4), 23, False),
    ]
    out = Path(__file__).with_name("golden_v1.json")
    out.write_text(json.dumps(g, indent=1) + "\n")
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
