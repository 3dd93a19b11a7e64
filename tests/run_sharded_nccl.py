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
    cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(9, 2, 1 << log_n)], alpha, delta)
    gc = [pkg.AirPermutationConfig(c.a_columns_ids, c.b_columns_ids, c.b_inverse_id, c.check_id) for c in cfgs]
    fri = pkg.FriConfig(num_queries=21)
    sharded = pkg.prove_sharded(comm, fri, gc, trace, [alpha, delta])
    single = pkg.prove(ctx, fri, gc, trace, [alpha, delta])
    ok = np.array_equal(sharded.words, single.words)
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
