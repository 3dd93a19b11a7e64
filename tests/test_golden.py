"""Golden fixtures (tests/golden/golden_v1.json, made by tests/golden/make_golden.py).

CPU half: the Python oracle and the C port reproduce the frozen vectors.  The GPU half
(tests/test_gpu_golden.py) holds the CUDA library to the same file.  The fixtures freeze the
oracle -- the reference itself ships no known-answer vectors."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import air as OA
from oracle import cport
from oracle import dft as OD
from oracle import field as F
from oracle import merkle as OM
from oracle import poseidon2 as OP
from oracle import stark as OS
from oracle import trace as OT
from tests.proofs import flat_from_dict

GOLD = json.loads((Path(__file__).parent / "golden" / "golden_v1.json").read_text())


def ix(s):
    return int(s, 16)


def test_constants_match_arkworks_published_values():
    # ark-bls12-377 0.5.0 `FrConfig`: MODULUS, GENERATOR = 22, TWO_ADICITY = 47, TWO_ADIC_ROOT_OF_UNITY
    assert ix(GOLD["modulus"]) == F.R_MOD == 8444461749428370424248824938781546531375899335154063827935233455917409239041
    assert ix(GOLD["generator"]) == 22
    root = ix(GOLD["two_adic_root_2_47"])
    assert root == 8065159656716812877374967518403273466521432693661810619979959746626482506078
    assert pow(root, 1 << 47, F.R_MOD) == 1 and pow(root, 1 << 46, F.R_MOD) == F.R_MOD - 1
    assert ix(GOLD["mont_r"]) == (1 << 256) % F.R_MOD and ix(GOLD["mont_r2"]) == pow(2, 512, F.R_MOD)


def test_field_vectors_python_and_c():
    a = [ix(v["a"]) for v in GOLD["field"]]
    b = [ix(v["b"]) for v in GOLD["field"]]
    for v, x, y in zip(GOLD["field"], a, b):
        assert F.add(x, y) == ix(v["add"]) and F.sub(x, y) == ix(v["sub"]) and F.mul(x, y) == ix(v["mul"])
        assert (F.inv(x) if x else 0) == ix(v["inv_a"]) and F.halve(x) == ix(v["halve_a"])
    assert cport.fr_mul(a, b) == [ix(v["mul"]) for v in GOLD["field"]]


@pytest.mark.parametrize("entry", GOLD["poseidon2"], ids=lambda e: f"d{e['sbox_d']}")
def test_poseidon2_vectors_python_and_c(entry):
    p = OP.Poseidon2Params.from_seed(entry["seed"], sbox_d=entry["sbox_d"], rounds_f=entry["rounds_f"], rounds_p=entry["rounds_p"])
    ins = [[ix(v) for v in c["in"]] for c in entry["cases"]]
    outs = [[ix(v) for v in c["out"]] for c in entry["cases"]]
    assert [OP.permute(p, s) for s in ins] == outs
    cport.set_poseidon2(p)
    assert cport.permute(ins) == outs


def test_sponge_lde_merkle_vectors():
    p = OP.Poseidon2Params.from_seed(0xB200, sbox_d=5)
    for c in GOLD["sponge"]:
        assert OP.hash_iter(p, [ix(v) for v in c["row"]]) == ix(c["digest"])
    mat = [[ix(v) for v in r] for r in GOLD["lde"]["in"]]
    lde = OD.coset_lde_batch(mat, GOLD["lde"]["added_bits"], ix(GOLD["lde"]["shift"]))
    assert lde == [[ix(v) for v in r] for r in GOLD["lde"]["out_bitrev_storage"]]
    assert lde == OD.coset_lde_batch_naive(mat, GOLD["lde"]["added_bits"], ix(GOLD["lde"]["shift"]))
    tree = OM.MerkleTree(p, [lde])
    assert tree.root == ix(GOLD["merkle"]["root"])
    assert tree.layers == [[ix(v) for v in l] for l in GOLD["merkle"]["layers"]]
    assert tree.open_batch(5)[1] == [ix(v) for v in GOLD["merkle"]["open_5"]["siblings"]]


def instance(case):
    rng = F.SplitMix64(case["seed"])
    alpha, delta = rng.next_fr(), rng.next_fr()
    assert alpha == ix(case["alpha"]) and delta == ix(case["delta"])
    cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(case["seed"], case["cols"], 1 << case["log_n"])], alpha, delta)
    return cfgs, trace, [alpha, delta]


def check_flat(case, words):
    words = np.asarray(words, dtype=np.uint64)
    assert words.size == case["n_words"]
    assert hashlib.sha256(words.astype("<u8").tobytes()).hexdigest() == case["sha256_flat_le_u64"]
    if "flat_words_hex" in case:
        assert [int(w) for w in words] == [ix(v) for v in case["flat_words_hex"]]


@pytest.mark.parametrize("case", GOLD["proofs"], ids=lambda c: f"2^{c['log_n']}x{c['cols']}")
def test_proof_vectors_python_and_c(case):
    p = OP.Poseidon2Params.from_seed(0xB200, sbox_d=5)
    cfgs, trace, publics = instance(case)
    fri = OS.FriConfig(**case["fri"])
    cport.set_poseidon2(p)
    check_flat(case, cport.prove(fri, cfgs, trace, publics))
    if case["log_n"] <= 4:   # the Python prover too (seconds)
        dbg = {}
        proof = OS.prove(p, fri, cfgs, trace, publics, dbg)
        assert proof["commitments"]["trace"] == ix(case["trace_commit"])
        assert proof["commitments"]["quotient_chunks"] == ix(case["quotient_commit"])
        assert list(dbg["query_indices"]) == case["query_indices"]
        check_flat(case, flat_from_dict(proof, dbg["query_indices"]))


# ---- frozen rejection reasons of the verifier (tests/golden/golden_verify_v1.json) ------------------------------
VGOLD = json.loads((Path(__file__).parent / "golden" / "golden_verify_v1.json").read_text())


def verify_fixture():
    """(case, cfgs, publics, words, vectors) of the verify fixture; vectors are (word, bit, expected code)."""
    case = GOLD["proofs"][VGOLD["proof"]]
    cfgs, _, publics = instance(case)
    words = np.array([ix(v) for v in case["flat_words_hex"]], dtype=np.uint64)
    return case, cfgs, publics, words, [tuple(v) for v in VGOLD["vectors"]]


def test_c_port_reproduces_the_frozen_rejection_reasons(p2params):
    from oracle import cport
    cport.set_poseidon2(p2params)
    case, cfgs, publics, words, vectors = verify_fixture()
    fri = OS.FriConfig(**case["fri"])
    pub = np.array([F.to_mont_limbs(x) for x in publics], dtype=np.uint64)
    w = OA.air_width(cfgs)
    assert cport.verify_limbs(fri, case["log_n"], w, cfgs, pub, words) == 0
    assert len(vectors) >= 40 and {c for _, _, c in vectors} >= {1, 2, 3, 4}
    for word, bit, code in vectors:
        bad = words.copy()
        bad[word] ^= np.uint64(1 << bit)
        assert cport.verify_limbs(fri, case["log_n"], w, cfgs, pub, bad) == code, (word, bit)


def _round2():
    import json
    from pathlib import Path
    return json.loads((Path(__file__).parent / "golden" / "golden_round2_v1.json").read_text())


def _fnv(words):
    h = 0xcbf29ce484222325
    for w in words.tolist():
        h = ((h ^ w) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def test_round2_fixture_fork_parameters_and_serialised_proof(pkg, p2params):
    """golden_round2_v1.json: the C port reproduces the proof hash under every non-default fork-only parameter, and the
    library's serialiser writes the committed bytes of the default-parameter proof (and reads them back)."""
    import numpy as np
    from oracle import air as OA
    from oracle import cport
    from oracle import field as F
    from oracle import stark as OS
    from tests.test_oracle_params import small_case
    g = _round2()
    cport.set_poseidon2(p2params)
    fri = OS.FriConfig(**g["fri"])
    cfgs, trace, publics = small_case(seed=31, log_n=4, c=2)
    try:
        for name, v in g["params"].items():
            cport.set_field_consts(int(v["generator"], 16), int(v["two_adic_root"], 16))
            cport.set_transcript_flags(v["alpha_before_openings"], v["observe_opened_values"])
            words = cport.prove(fri, cfgs, trace, publics)
            assert _fnv(words) == v["fnv1a64"], name
            assert F.from_mont_limbs(words[:4]) == int(v["trace_commit"], 16), name
            if name == "default":
                proof = pkg.Proof(words, 4, OA.air_width(cfgs), OA.log_quotient_degree(cfgs), pkg.FriConfig(**g["fri"]))
                assert proof.serialize().hex() == g["serialized_hex"]
                assert pkg.Proof.deserialize(bytes.fromhex(g["serialized_hex"])).to_dict()[0] == proof.to_dict()[0]
    finally:
        cport.set_field_consts(F.GENERATOR, F.TWO_ADIC_ROOT)
        cport.set_transcript_flags()
