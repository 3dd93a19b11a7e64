"""linea-stark-prover_b200: B200-native (sm_100a) proving backend for the Plonky3
prover loop of distributed-lab/linea-stark-prover (permutation AIR over BLS12-377 Fr).

The directory name carries a hyphen (it mirrors the reference's repo name), so
it is imported under the module name `linea_stark_prover_b200` via
`__graft_entry__.load_package()`.
"""
from . import ffi  # noqa: F401
from .backend import (AirLookupConfig, AirPermutationConfig, BackendError, Comm, Context, FriConfig, GpuDft, GpuMmcs, Mat, Proof, Tree,  # noqa: F401
                      eval_at, reduce_openings, fri_fold, from_mont_array, prove, prove_sharded, quotient_air, quotient_permutation, read_raw_lookup_trace, read_raw_permutation_trace, read_lookup_trace_once, read_permutation_trace_once,
                      shard_plan, to_mont_array, VerificationError, verify, verify_code, VERIFY_REASONS)
