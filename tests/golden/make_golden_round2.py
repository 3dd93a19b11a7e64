#!/usr/bin/env python
"""Generates tests/golden/golden_round2_v1.json from the Python big-integer oracle (run in the build container):

    python tests/golden/make_golden_round2.py

Freezes what round 2 added, so that the C port, the CUDA library and later edits of the oracle are held to the same
numbers: (1) one small proof under every NON-DEFAULT value of the fork-only parameters (coset shift GENERATOR,
two_adic_generator(47), transcript order of Pcs::open) -- the proof's 64-bit FNV-1a hash and its first commitment;
(2) the serialised form (lsp_proof_serialize's byte format) of the default-parameter proof, written by the independent
field-order writer of tests/test_proof_serialization.py.  Like golden_v1.json these vectors come from the oracle, NOT from
the Rust reference (which cannot be built or imported here and ships no vectors: DESIGN.md section 2)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import air as OA  # noqa: E402
from oracle import field as F  # noqa: E402
from oracle import poseidon2 as OP  # noqa: E402
from oracle import stark as OS  # noqa: E402
from tests.proofs import flat_from_dict  # noqa: E402
from tests.test_oracle_params import CASES, small_case  # noqa: E402
from tests.test_proof_serialization import write_reference_shape  # noqa: E402

FRI = dict(log_blowup=2, log_final_poly_len=1, num_queries=5, proof_of_work_bits=2)


def fnv1a64(words):
    h = 0xcbf29ce484222325
    for w in np.asarray(words, dtype=np.uint64).tolist():
        h = ((h ^ w) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def main():
    p = OP.Poseidon2Params.from_seed(0xB200, sbox_d=5)
    fri = OS.FriConfig(**FRI)
    cfgs, trace, publics = small_case(seed=31, log_n=4, c=2)
    out = {"note": "oracle-generated (see make_golden_round2.py); NOT reference vectors", "fri": FRI, "case": "small_case(seed=31, log_n=4, c=2)",
           "params": {}}
    for name, case in [("default", {})] + list(CASES.items()):
        consts, flags = case.get("consts", (F.GENERATOR, F.TWO_ADIC_ROOT)), case.get("flags", (True, False))
        F.set_field_consts(*consts)
        OS.set_transcript_flags(*flags)
        dbg = {}
        proof = OS.prove(p, fri, cfgs, trace, publics, dbg)
        OS.verify(p, fri, cfgs, proof, publics)
        words = flat_from_dict(proof, dbg["query_indices"])
        out["params"][name] = {"generator": format(consts[0], "x"), "two_adic_root": format(consts[1], "x"), "alpha_before_openings": flags[0],
                               "observe_opened_values": flags[1], "fnv1a64": fnv1a64(words), "trace_commit": format(proof["commitments"]["trace"], "x")}
        if name == "default":
            out["serialized_hex"] = write_reference_shape(proof, fri, OA.air_width(cfgs), OA.log_quotient_degree(cfgs)).hex()
    F.set_field_consts()
    OS.set_transcript_flags()
    path = Path(__file__).with_name("golden_round2_v1.json")
    path.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
