"""One small invocation of the hot path on cuda:0, checked against the oracle
(used by __graft_entry__.smoke())."""
import __graft_entry__ as g
from oracle import air as OA
from oracle import field as F
from oracle import stark as OS
from oracle import trace as OT
from oracle.poseidon2 import Poseidon2Params


def run_smoke(log_n: int = 6, c: int = 3):
    pkg = g.load_package()
    p = Poseidon2Params.from_seed(0xB200, sbox_d=5)
    ctx = pkg.Context(0)  # raises without a CUDA device / library: no CPU fallback
    ctx.set_poseidon2(p.sbox_d, p.rounds_f, p.rounds_p, p.flat_constants(), p.internal_diag_m1)
    rng = F.SplitMix64(11)
    alpha, delta = rng.next_fr(), rng.next_fr()
    cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(5, c, 1 << log_n)], alpha, delta)
    fri = OS.FriConfig()
    dbg = {}
    oproof = OS.prove(p, fri, cfgs, trace, [alpha, delta], dbg)
    gcfgs = [pkg.AirPermutationConfig(x.a_columns_ids, x.b_columns_ids, x.b_inverse_id, x.check_id) for x in cfgs]
    launches0 = ctx.kernel_launches()
    gproof = pkg.prove(ctx, pkg.FriConfig(), gcfgs, trace, [alpha, delta])
    gd, idx = gproof.to_dict()
    assert ctx.kernel_launches() > launches0, "no CUDA kernels were launched"
    assert gd == oproof and idx == dbg["query_indices"], "GPU proof differs from the oracle's"
    OS.verify(p, fri, cfgs, gd, [alpha, delta])
    pkg.verify(ctx, pkg.FriConfig(), gcfgs, gproof, [alpha, delta])      # the device verifier agrees
    bad = gproof.words.copy()
    bad[4 * 2] ^= 1
    assert pkg.verify_code(ctx, pkg.FriConfig(), gcfgs, bad, [alpha, delta], log_n, gproof.width) != 0
    assert OA.check_constraints(cfgs, trace, [alpha, delta])
    ctx.close()
    print(f"smoke ok: 2^{log_n}-row permutation AIR proved on cuda:0, bit-exact with the oracle, verifier accepts")
