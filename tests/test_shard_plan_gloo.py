"""The N>1 host logic on CPU: world_size-2 gloo processes each take their shard of the LDE
(shard_plan), build their Merkle subtree and fold their FRI slice with the ORACLE, exchange
subtree roots with all_gather, and must reproduce the unsharded commitment and fold."""
import os
import sys
from pathlib import Path

import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    from oracle import dft as OD
    from oracle import field as F
    from oracle import merkle as OM
    from oracle import poseidon2 as OP
    from oracle import stark as OS
    pkg = g.load_package()
    p = OP.Poseidon2Params.from_seed(1)
    log_n, blow = 3, 2
    rng = F.SplitMix64(4)
    mat = [[rng.next_fr() for _ in range(3)] for _ in range(1 << log_n)]
    plan = pkg.shard_plan(log_n, blow, world, rank)
    # this rank's rows of the bit-reversed LDE = its cosets, each a size-N coset evaluation
    n = 1 << log_n
    w_l = F.two_adic_generator(log_n + blow)
    local = []
    for c in plan["cosets"]:
        block = OD.coset_lde_batch(mat, 0, F.GENERATOR * pow(w_l, c, F.R_MOD) % F.R_MOD)
        local += block
    full = OD.coset_lde_batch(mat, blow, F.GENERATOR)
    assert local == full[plan["row0"]:plan["row0"] + plan["rows"]]
    sub = OM.MerkleTree(p, [local])
    roots = [None] * world
    dist.all_gather_object(roots, sub.root)
    top = roots
    while len(top) > 1:
        top = [OP.compress(p, top[2 * i], top[2 * i + 1]) for i in range(len(top) // 2)]
    assert top[0] == OM.MerkleTree(p, [full]).root
    # FRI fold of the local slice equals the slice of the global fold
    vec = [r[0] for r in full]
    beta = rng.next_fr()
    gfold = OS.fold_matrix(beta, [[vec[2 * j], vec[2 * j + 1]] for j in range(len(vec) // 2)])
    h = len(vec) // 2
    log_h = h.bit_length() - 1
    g_inv = F.inv(F.two_adic_generator(log_h + 1))
    j0 = plan["row0"] // 2
    mine = []
    for j in range(plan["rows"] // 2):
        t = F.halve(beta) * pow(g_inv, F.reverse_bits_len(j0 + j, log_h), F.R_MOD) % F.R_MOD
        lo, hi = vec[2 * (j0 + j)], vec[2 * (j0 + j) + 1]
        mine.append((F.halve((lo + hi) % F.R_MOD) + t * (lo - hi)) % F.R_MOD)
    assert mine == gfold[j0:j0 + plan["rows"] // 2]
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def _worker_fraction(rank, world, port, q):
    """More ranks than cosets (blowup 1, two ranks): each rank owns HALF of the one coset.  Its rows are a sub-coset: the
    trace polynomial folded to half its degree (x^M is constant there) and evaluated on sigma * H_M; its next rows are its
    ring neighbour's rows; the subtree roots still assemble the commitment."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    from oracle import dft as OD
    from oracle import field as F
    from oracle import merkle as OM
    from oracle import poseidon2 as OP
    pkg = g.load_package()
    p = OP.Poseidon2Params.from_seed(1)
    log_n, blow = 4, 0
    n = 1 << log_n
    rng = F.SplitMix64(9)
    mat = [[rng.next_fr() for _ in range(2)] for _ in range(n)]
    plan = pkg.shard_plan(log_n, blow, world, rank)
    k0, s = plan["fraction"]
    assert s == world and plan["rows"] == n // s and plan["cosets"] == [0]
    m = n // s
    full = OD.coset_lde_batch(mat, blow, F.GENERATOR)                      # bit-reversed rows of p(g w_N^k)
    coeffs = [OD.idft([row[c] for row in mat]) for c in range(2)]          # natural-order coefficients of each column
    w_n = F.two_adic_generator(log_n)
    sigma = F.GENERATOR * pow(w_n, k0, F.R_MOD) % F.R_MOD
    u = pow(sigma, m, F.R_MOD)
    log_m = m.bit_length() - 1
    mine = [[0, 0] for _ in range(m)]
    for c in range(2):
        a = coeffs[c]
        folded = [sum(a[i0 + m * t] * pow(u, t, F.R_MOD) for t in range(s)) % F.R_MOD for i0 in range(m)]
        w_m = F.two_adic_generator(log_m)
        for mm in range(m):
            x = sigma * pow(w_m, mm, F.R_MOD) % F.R_MOD
            mine[F.reverse_bits_len(mm, log_m)][c] = sum(f * pow(x, i, F.R_MOD) for i, f in enumerate(folded)) % F.R_MOD
    assert mine == full[plan["row0"]:plan["row0"] + plan["rows"]]
    # The quotient pairs every row with its NEXT trace row: those are the rows of ONE neighbour, the rank of residue
    # k0 + 1 -- at the same local row, or at the row of the following point when the residue wrapped to 0.  One ring step.
    everyone = [None] * world
    dist.all_gather_object(everyone, (k0, mine))
    theirs = dict(everyone)[(k0 + 1) % s]
    rotated = k0 == s - 1
    for r in range(m):
        k = F.reverse_bits_len(plan["row0"] + r, log_n)
        want = full[F.reverse_bits_len((k + 1) % n, log_n)]
        rr = F.reverse_bits_len((F.reverse_bits_len(r, log_m) + 1) % m, log_m) if rotated else r
        assert theirs[rr] == want
    sub = OM.MerkleTree(p, [mine])
    roots = [None] * world
    dist.all_gather_object(roots, sub.root)
    top = roots
    while len(top) > 1:
        top = [OP.compress(p, top[2 * i], top[2 * i + 1]) for i in range(len(top) // 2)]
    assert top[0] == OM.MerkleTree(p, [full]).root
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def test_shard_plan_fraction_of_a_coset_world2_gloo():
    world, port = 2, 29537
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_fraction, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    assert sorted(q.get() for _ in range(world)) == [0, 1]


def test_shard_plan_world2_gloo():
    world, port = 2, 29531
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    assert sorted(q.get() for _ in range(world)) == [0, 1]


def test_shard_plan_shapes(pkg):
    import pytest
    assert pkg.shard_plan(19, 3, 8, 3) == dict(row0=3 << 19, rows=1 << 19, blocks=[3], cosets=[6], fraction=(0, 1))
    assert pkg.shard_plan(4, 3, 2, 1)["cosets"] == [1, 5, 3, 7]
    # more ranks than cosets (BASELINE configs[3]: blowup 2 on 8 GPUs): rank 5 owns quarter 1 of block 1 = coset 1, the
    # trace rows k = 2 (mod 4) of it
    assert pkg.shard_plan(20, 1, 8, 5) == dict(row0=5 << 18, rows=1 << 18, blocks=[1], cosets=[1], fraction=(2, 4))
    with pytest.raises(pkg.BackendError):
        pkg.shard_plan(3, 1, 8, 0)       # fewer than 8 rows per rank
    with pytest.raises(pkg.BackendError):
        pkg.shard_plan(4, 1, 3, 0)
