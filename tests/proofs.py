"""Proof (dict, oracle/stark.py shape) <-> flat word layout shared by the CUDA library
(linea-stark-prover_b200/host/prover.cu) and the C oracle."""
import numpy as np

from oracle.field import to_mont_limbs


def flat_from_dict(proof: dict, indices) -> np.ndarray:
    out = []
    put = lambda v: out.append(to_mont_limbs(v))
    put(proof["commitments"]["trace"])
    put(proof["commitments"]["quotient_chunks"])
    ov = proof["opened_values"]
    for v in ov["trace_local"] + ov["trace_next"]:
        put(v)
    for ch in ov["quotient_chunks"]:
        put(ch[0])
    fp = proof["opening_proof"]
    for v in fp["commit_phase_commits"] + fp["final_poly"]:
        put(v)
    put(fp["pow_witness"])
    for qp, idx in zip(fp["query_proofs"], indices):
        out.append([idx, 0, 0, 0])
        for bo in qp["input_proof"]:
            for row in bo["opened_values"]:
                for v in row:
                    put(v)
            for v in bo["opening_proof"]:
                put(v)
        for st in qp["commit_phase_openings"]:
            put(st["sibling_value"])
            for v in st["opening_proof"]:
                put(v)
    return np.array(out, dtype=np.uint64).reshape(-1)
