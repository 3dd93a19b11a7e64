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


def dict_from_flat(words, log_n: int, width: int, log_q: int, fri):
    """Inverse of flat_from_dict: (proof dict with canonical integers, stored query indices)."""
    from oracle.field import from_mont_limbs
    raw = np.asarray(words, dtype=np.uint64).reshape(-1, 4)
    q = 1 << log_q
    log_l = log_n + fri.log_blowup
    rounds = log_n - fri.log_final_poly_len
    f = 1 << (fri.log_blowup + fri.log_final_poly_len)
    pos = 0

    def take(k):
        nonlocal pos
        out = [from_mont_limbs(r) for r in raw[pos:pos + k]]
        pos += k
        return out

    trace_commit, quot_commit = take(2)
    local, nxt = take(width), take(width)
    chunks = [[x] for x in take(q)]
    commits, final_poly = take(rounds), take(f)
    pow_witness = take(1)[0]
    queries, indices = [], []
    for _ in range(fri.num_queries):
        indices.append(int(raw[pos][0]) if not raw[pos][1:].any() else -1)
        pos += 1
        row = take(width)
        ip = [dict(opened_values=[row], opening_proof=take(log_l))]
        row = take(q)
        ip.append(dict(opened_values=[[x] for x in row], opening_proof=take(log_l)))
        steps = []
        for r in range(rounds):
            sib = take(1)[0]
            steps.append(dict(sibling_value=sib, opening_proof=take(log_l - 1 - r)))
        queries.append(dict(input_proof=ip, commit_phase_openings=steps))
    assert pos == len(raw)
    return dict(commitments=dict(trace=trace_commit, quotient_chunks=quot_commit),
                opened_values=dict(trace_local=local, trace_next=nxt, quotient_chunks=chunks),
                opening_proof=dict(commit_phase_commits=commits, query_proofs=queries, final_poly=final_poly,
                                   pow_witness=pow_witness),
                degree_bits=log_n), indices


def fast_cbor_permutation_trace(be: "np.ndarray", n: int, c: int, name: str) -> bytes:
    """`RawPermutationTrace` CBOR (serde layout: map{a, b, name}, columns of rows of 32-tuples of small ints) from the
    row-major raw bytes be[n, 2c, 32] -- vectorised with numpy so that a 2^19-row file (hundreds of MB) takes seconds;
    byte-identical to `oracle.trace.encode_raw_permutation_trace` (tests/test_cbor_input.py checks that)."""
    import numpy as np

    def head(major, v):
        if v < 24:
            return bytes([major << 5 | v])
        for info, size in ((24, 1), (25, 2), (26, 4), (27, 8)):
            if v < 1 << (8 * size):
                return bytes([major << 5 | info]) + int(v).to_bytes(size, "big")

    def column(col):                                  # col: uint8[n, 32]
        ext = np.empty((n, 34), dtype=np.uint8)       # element head 0x98 0x20, then the 32 value bytes
        ext[:, 0], ext[:, 1] = 0x98, 0x20
        ext[:, 2:] = col
        flat = ext.reshape(-1)
        big = flat >= 24
        big.reshape(n, 34)[:, :2] = False             # the head bytes are written verbatim
        pos = np.arange(flat.size, dtype=np.int64) + np.concatenate(([0], np.cumsum(big)[:-1]))
        out = np.empty(flat.size + int(big.sum()), dtype=np.uint8)
        out[pos + big] = flat
        out[pos[big]] = 0x18                          # uint8 >= 24 is 0x18 xx
        return head(4, n) + out.tobytes()

    be = np.ascontiguousarray(be, dtype=np.uint8).reshape(n, 2 * c, 32)
    side = lambda j0: head(4, c) + b"".join(column(be[:, j0 + j, :]) for j in range(c))
    text = lambda t: head(3, len(t)) + t.encode()
    return head(5, 3) + text("a") + side(0) + text("b") + side(c) + text("name") + text(name)
