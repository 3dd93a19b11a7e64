"""The only run-time artefact the reference ships for this path is its own `bench.log` (a `tracing-forest` dump of one
prove + verify, README.md:19-21).  It holds no field values, but it does pin SHAPES -- matrix dimensions, how many
times each PCS step runs, the length of the final polynomial -- and the oracle (and through it the CUDA path, whose
proofs must equal the oracle's) is held to every one of them here.  The numbers are quoted from bench.log with their
line numbers; when the reference tree is present they are re-read from the file as well."""
import re
from pathlib import Path

import pytest

from oracle import air as OA
from oracle import field as F
from oracle import stark as OS
from oracle import trace as OT

BENCH_LOG = Path("/root/reference/bench.log")

# bench.log:2-13   6 'from' + 6 'to' columns of length 524288
# bench.log:20     coset_lde_batch dims: 14x524288 | added_bits: 3            (trace: 2*6 + 2 columns)
# bench.log:23-30  eight coset_lde_batch dims: 1x524288 | added_bits: 3       (one per quotient chunk, each committed alone)
# bench.log:33,36  two  `reduce matrix quotient` dims: 14x4194304             (trace opened at zeta and zeta*g)
# bench.log:39-60  eight `reduce matrix quotient` dims: 1x4194304             (each chunk opened at zeta only)
# bench.log:65     divide_by_height dims: 1x8                                 (final polynomial: 2^(3+0) values)
QUOTED = dict(cols_per_side=6, rows=524288, trace_dims=(14, 524288), added_bits=3, chunk_ldes=8, chunk_dims=(1, 524288),
              trace_reductions=2, lde_rows=4194304, chunk_reductions=8, final_poly_len=8)


def test_quoted_numbers_are_what_bench_log_says():
    if not BENCH_LOG.exists():
        pytest.skip("reference tree not present (GPU box): the quoted constants stand on their own")
    lines = BENCH_LOG.read_text().splitlines()
    assert sum("Found 'from' column" in l for l in lines) == QUOTED["cols_per_side"] == sum("Found 'to' column" in l for l in lines)
    assert all(f"length: {QUOTED['rows']}" in l for l in lines if "Found '" in l and "column" in l)
    ldes = [tuple(map(int, re.search(r"dims: (\d+)x(\d+) \| added_bits: (\d+)", l).groups())) for l in lines if "coset_lde_batch" in l]
    assert ldes[0] == (*QUOTED["trace_dims"], QUOTED["added_bits"])
    assert ldes[1:] == [(*QUOTED["chunk_dims"], QUOTED["added_bits"])] * QUOTED["chunk_ldes"]
    red = [tuple(map(int, re.search(r"dims: (\d+)x(\d+)", l).groups())) for l in lines if "reduce matrix quotient" in l]
    assert red == [(14, QUOTED["lde_rows"])] * QUOTED["trace_reductions"] + [(1, QUOTED["lde_rows"])] * QUOTED["chunk_reductions"]
    fin = [l for l in lines if "divide_by_height" in l]
    assert len(fin) == 1 and f"dims: 1x{QUOTED['final_poly_len']}" in fin[0]
    order = [k for l in lines for k in ("commit to trace data", "compute quotient polynomial", "commit to quotient poly chunks",
                                        "open [", "compute_inverse_denominators", "FRI prover", "commit phase", "grind for proof-of-work",
                                        "query phase", "verify [") if k in l]
    assert order == ["commit to trace data", "compute quotient polynomial", "commit to quotient poly chunks", "open [",
                     "compute_inverse_denominators", "FRI prover", "commit phase", "grind for proof-of-work", "query phase", "verify ["]


def test_oracle_has_the_shapes_of_the_reference_run(p2params):
    """The same AIR shape (6 + 6 columns) and FRI parameters (main.rs:58-64) at a height the Python oracle can prove:
    every count that does not depend on the height must equal the log's, the others must scale with it."""
    c, log_n = QUOTED["cols_per_side"], 4
    cfg = OA.AirPermutationConfig.standard(c)
    assert OA.air_width([cfg]) == QUOTED["trace_dims"][0] == 2 * c + 2            # air/src/air_permutation.rs:21-23
    fri = OS.FriConfig(log_blowup=QUOTED["added_bits"], log_final_poly_len=0, num_queries=33, proof_of_work_bits=0)
    rng = F.SplitMix64(2024)
    alpha, delta = rng.next_fr(), rng.next_fr()
    cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(7, c, 1 << log_n)], alpha, delta)
    assert (len(trace[0]), len(trace)) == (QUOTED["trace_dims"][0], 1 << log_n)
    dbg = {}
    proof = OS.prove(p2params, fri, cfgs, trace, [alpha, delta], dbg)
    # LDE: rows x 2^added_bits, all 14 columns in one matrix (bench.log:20,33)
    assert len(dbg["trace_lde"]) == (1 << log_n) << QUOTED["added_bits"] and len(dbg["trace_lde"][0]) == QUOTED["trace_dims"][0]
    assert QUOTED["lde_rows"] == QUOTED["rows"] << QUOTED["added_bits"]
    # quotient chunks: each its own 1-column matrix of the trace's height, LDE'd separately (bench.log:23-30).  The log's
    # run had 8 of them (an older AIR of higher degree); HEAD's permutation AIR has degree 3 => 2 (caveat ii)
    q = 1 << OA.log_quotient_degree(cfgs)
    assert q == 2 and len(dbg["quotient_ldes"]) == q
    assert all(len(m) == len(dbg["trace_lde"]) and len(m[0]) == QUOTED["chunk_dims"][0] for m in dbg["quotient_ldes"])
    # openings: the trace at two points, every chunk at one (bench.log:33-60)
    ov = proof["opened_values"]
    assert len(ov["trace_local"]) == len(ov["trace_next"]) == QUOTED["trace_dims"][0]
    assert len(ov["quotient_chunks"]) == q and all(len(ch) == QUOTED["chunk_dims"][0] for ch in ov["quotient_chunks"])
    # FRI: final polynomial of 2^(log_blowup + log_final_poly_len) = 8 values (bench.log:65), log2(height) folding rounds
    fp = proof["opening_proof"]
    assert len(fp["final_poly"]) == QUOTED["final_poly_len"]
    assert len(fp["commit_phase_commits"]) == log_n and len(fp["query_proofs"]) == 33
    OS.verify(p2params, fri, cfgs, proof, [alpha, delta])


def test_library_stage_names_follow_the_reference_spans(pkg):
    """`timings_ms_out` of lsp_prove_* is indexed by the reference's span tree (bench.log:18-67), in its order."""
    assert pkg.backend.STAGE_NAMES == ["commit_trace_lde", "commit_trace_merkle", "quotient", "commit_quotient", "open_reduce",
                                       "fri_commit_phase", "grind_query", "d2h"]
    # proof size of the reference's run shape through the library's own formula: 19 rounds, 8 final values
    import ctypes as C
    lib = pkg.ffi.load()
    fri = pkg.FriConfig().c_struct()
    w, log_n, log_l, q = 14, 19, 22, 2
    per_query = 1 + (w + log_l) + (q + log_l) + sum(1 + (log_l - 1 - r) for r in range(log_n))
    assert lib.lsp_proof_words(log_n, w, 1, C.byref(fri)) == 4 * (2 + 2 * w + q + log_n + QUOTED["final_poly_len"] + 1 + 33 * per_query)
