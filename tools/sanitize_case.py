"""Small end-to-end case for compute-sanitizer: permutation prove, lookup prove, sharded (local communicator)
prove, CBOR witness -- each checked against the oracle."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as g
from oracle import air as OA, field as F, stark as OS, trace as OT
from oracle.poseidon2 import Poseidon2Params

pkg = g.load_package()
p = Poseidon2Params.from_seed(0xB200, sbox_d=5)
ctx = pkg.Context(0)
ctx.set_poseidon2(p.sbox_d, p.rounds_f, p.rounds_p, p.flat_constants(), p.internal_diag_m1)
rng = F.SplitMix64(1)
alpha, delta = rng.next_fr(), rng.next_fr()


def gcfgs(cfgs):
    out = []
    for c in cfgs:
        if isinstance(c, OA.AirLookupConfig):
            out.append(pkg.AirLookupConfig(c.a_columns_ids, c.b_columns_ids, c.a_filter_id, c.b_filter_id, c.a_inverses_id,
                                           c.b_inverses_id, c.occurrences_id, c.check_id))
        else:
            out.append(pkg.AirPermutationConfig(c.a_columns_ids, c.b_columns_ids, c.b_inverse_id, c.check_id))
    return out


fri = dict(log_blowup=3, log_final_poly_len=0, num_queries=5, proof_of_work_bits=2)
cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(2, 3, 64)], alpha, delta)
gd, _ = pkg.prove(ctx, pkg.FriConfig(**fri), gcfgs(cfgs), trace, [alpha, delta]).to_dict()
assert gd == OS.prove(p, OS.FriConfig(**fri), cfgs, trace, [alpha, delta])
comm = pkg.Comm.local(ctx, 4)
sd, _ = pkg.prove_sharded(comm, pkg.FriConfig(**fri), gcfgs(cfgs), trace, [alpha, delta]).to_dict()
comm.close()
assert sd == gd
# more ranks than cosets: 8 ranks on 2 cosets (fold + sub-coset NTT, next-row LDE, summed chunk shares)
fri1 = dict(fri, log_blowup=1)
one = pkg.prove(ctx, pkg.FriConfig(**fri1), gcfgs(cfgs), trace, [alpha, delta])
comm = pkg.Comm.local(ctx, 8)
assert (pkg.prove_sharded(comm, pkg.FriConfig(**fri1), gcfgs(cfgs), trace, [alpha, delta]).words == one.words).all()
comm.close()
# verifier, serialisation round trip, Pcs::open pieces, other field parameters
pkg.verify(ctx, pkg.FriConfig(**fri1), gcfgs(cfgs), pkg.Proof.deserialize(one.serialize()), [alpha, delta])
lde, co = pkg.GpuDft(ctx).coset_lde_batch(ctx.upload(trace), 1, F.GENERATOR, want_coeffs=True)
ys = pkg.eval_at(ctx, co, 12345)
pkg.reduce_openings(ctx, [(lde, 12345, ys), (lde, 999, pkg.eval_at(ctx, co, 999))], 77).rows()
ctx.set_field_consts(5, pow(F.TWO_ADIC_ROOT, 3, F.R_MOD))
ctx.set_transcript_flags(False, True)
alt = pkg.prove(ctx, pkg.FriConfig(**fri1), gcfgs(cfgs), trace, [alpha, delta])
pkg.verify(ctx, pkg.FriConfig(**fri1), gcfgs(cfgs), alt, [alpha, delta])
ctx.set_field_consts(F.GENERATOR, F.TWO_ADIC_ROOT)
ctx.set_transcript_flags()
cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(3, 1, 32)], alpha, delta, [OT.synthetic_lookup_input(4, 2, 2, 32, disabled_every=5)])
gd, _ = pkg.prove(ctx, pkg.FriConfig(**fri), gcfgs(cfgs), trace, [alpha, delta]).to_dict()
assert gd == OS.prove(p, OS.FriConfig(**fri), cfgs, trace, [alpha, delta])
a, b = OT.synthetic_permutation_input(5, 2, 128)
be, rows, nc, _ = pkg.read_raw_permutation_trace(OT.encode_raw_permutation_trace(a, b, "s"))
dev = ctx.permutation_trace_be(be, rows, nc, pkg.to_mont_array([alpha, delta]))
assert dev.rows() == OT.build_trace([(a, b)], alpha, delta)[1]
ctx.close()
print("sanitize case ok")
