"""torchrun entry: sharded prove over NCCL (one process per GPU) == single-GPU prove."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
from oracle import field as F  # noqa: E402
from oracle import trace as OT  # noqa: E402
from oracle.poseidon2 import Poseidon2Params  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    pkg = g.load_package()
    p = Poseidon2Params.from_seed(0xB200, sbox_d=5)
    ctx = pkg.Context(local)
    ctx.set_poseidon2(p.sbox_d, p.rounds_f, p.rounds_p, p.flat_constants(), p.internal_diag_m1)
    uid = [pkg.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = pkg.Comm.nccl(ctx, rank, world, uid[0])
    log_n = int(os.environ.get("LSP_LOG_N", "12"))
    rng = F.SplitMix64(3)
    alpha, delta = rng.next_fr(), rng.next_fr()
    oks = []
    # (1) the permutation AIR; (2) a lookup + permutation AIR of ODD width (15 + 4: the coefficient all-gather pads the matrix
    # to a multiple of the rank count), 4 quotient chunks; (3) blowup 2 -- with 4 or 8 ranks: more ranks than cosets (the
    # ring exchange of next rows and the summed chunk shares over NCCL)
    cases = [([OT.synthetic_permutation_input(9, 2, 1 << log_n)], [], dict(num_queries=21)),
             ([OT.synthetic_permutation_input(10, 1, 1 << (log_n - 2))], [OT.synthetic_lookup_input(11, 2, 2, 1 << (log_n - 2), disabled_every=7)],
              dict(num_queries=9)),
             ([OT.synthetic_permutation_input(12, 3, 1 << (log_n - 1))], [], dict(log_blowup=1, num_queries=13))]
    for perms, lookups, fri_kw in cases:
        cfgs, trace = OT.build_trace(perms, alpha, delta, lookups)
        gc = []
        for c in cfgs:
            if hasattr(c, "occurrences_id"):
                gc.append(pkg.AirLookupConfig(c.a_columns_ids, c.b_columns_ids, c.a_filter_id, c.b_filter_id, c.a_inverses_id, c.b_inverses_id,
                                              c.occurrences_id, c.check_id))
            else:
                gc.append(pkg.AirPermutationConfig(c.a_columns_ids, c.b_columns_ids, c.b_inverse_id, c.check_id))
        fri = pkg.FriConfig(**fri_kw)
        single = pkg.prove(ctx, fri, gc, trace, [alpha, delta])
        sharded = pkg.prove_sharded(comm, fri, gc, trace, [alpha, delta])                 # host trace: every rank uploads its rows
        dev = ctx.upload(trace)
        sharded_dev = pkg.prove_sharded(comm, fri, gc, dev, [alpha, delta])               # trace already on every device
        oks.append(np.array_equal(sharded.words, single.words) and np.array_equal(sharded_dev.words, single.words))
    ok = all(oks)
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    comm.close()
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        assert all(flags), flags
        print("sharded nccl ok", world, "ranks")


if __name__ == "__main__":
    main()
