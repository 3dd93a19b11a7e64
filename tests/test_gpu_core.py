"""GPU parity, first slice: Fr arithmetic (K1), Poseidon2 permutation / sponge (K4),
Merkle MMCS (K5) against the oracle -- all through the C ABI.  Bit-exact."""
import pytest

from oracle import field as F
from oracle import merkle as OM
from oracle import poseidon2 as OP

pytestmark = pytest.mark.gpu

EDGE = [0, 1, 2, F.R_MOD - 1, F.R_MOD - 2, F.MONT_R, F.MONT_R2, F.MONT_RINV, (1 << 252), (1 << 252) - 1,
        0xFFFFFFFF, 1 << 32, (1 << 64) - 1, 1 << 64, F.R_MOD // 2, F.R_MOD // 2 + 1]


def _rand(n, seed):
    rng = F.SplitMix64(seed)
    return [rng.next_fr() for _ in range(n)]


def test_fr_ops_random_and_edges(gctx):
    a = [x for x in EDGE for _ in EDGE] + _rand(20000, 1)
    b = [y for _ in EDGE for y in EDGE] + _rand(20000, 2)
    assert gctx.fr_op("add", a, b) == [(x + y) % F.R_MOD for x, y in zip(a, b)]
    assert gctx.fr_op("sub", a, b) == [(x - y) % F.R_MOD for x, y in zip(a, b)]
    assert gctx.fr_op("mul", a, b) == [x * y % F.R_MOD for x, y in zip(a, b)]
    assert gctx.fr_op("halve", a) == [F.halve(x) for x in a]


def test_fr_inverse(gctx):
    a = [x for x in EDGE if x] + _rand(2000, 3)
    got = gctx.fr_op("inv", a)
    assert got == [F.inv(x) for x in a]
    assert gctx.fr_op("inv", [0]) == [0]


def test_fr_op_empty(gctx):
    assert gctx.fr_op("mul", [], []) == []


@pytest.mark.parametrize("d", [3, 5, 7, 11, 17])
def test_poseidon2_permute_all_sbox_degrees(pkg, d):
    p = OP.Poseidon2Params.from_seed(100 + d, sbox_d=d)
    ctx = pkg.Context(0)
    ctx.set_poseidon2(d, p.rounds_f, p.rounds_p, p.flat_constants(), p.internal_diag_m1)
    rng = F.SplitMix64(7)
    states = [[0, 0, 0], [1, 2, 3], [F.R_MOD - 1] * 3] + [[rng.next_fr() for _ in range(3)] for _ in range(300)]
    assert ctx.permute(states) == [OP.permute(p, s) for s in states]
    ctx.close()


def test_poseidon2_generic_diag_and_round_counts(pkg):
    p = OP.Poseidon2Params.from_seed(9, sbox_d=5, rounds_f=6, rounds_p=11)
    p.internal_diag_m1 = (3, 5, 7)
    ctx = pkg.Context(0)
    ctx.set_poseidon2(5, 6, 11, p.flat_constants(), p.internal_diag_m1)
    states = [[i, i + 1, i + 2] for i in range(40)]
    assert ctx.permute(states) == [OP.permute(p, s) for s in states]
    ctx.close()


@pytest.mark.parametrize("w", [0, 1, 2, 3, 4, 7, 8, 14])
def test_sponge_widths(gctx, p2params, w):
    rng = F.SplitMix64(50 + w)
    rows = [[rng.next_fr() for _ in range(w)] for _ in range(33)]
    assert gctx.hash_rows(rows) == [OP.hash_iter(p2params, r) for r in rows]


def test_poseidon2_requires_constants(pkg):
    ctx = pkg.Context(0)
    with pytest.raises(pkg.BackendError):
        ctx.permute([[1, 2, 3]])
    with pytest.raises(pkg.BackendError):
        ctx.set_poseidon2(4, 8, 22, [0] * 46)  # unsupported S-box degree
    ctx.close()


@pytest.mark.parametrize("h,widths", [(1, [3]), (2, [1]), (8, [8]), (64, [2, 1]), (256, [1, 1]), (128, [14])])
def test_merkle_commit_open(pkg, gctx, p2params, h, widths):
    rng = F.SplitMix64(h * 31 + sum(widths))
    mats = [[[rng.next_fr() for _ in range(w)] for _ in range(h)] for w in widths]
    ot = OM.MerkleTree(p2params, mats)
    dm = [gctx.upload(m) for m in mats]
    assert all(d.rows() == m for d, m in zip(dm, mats))  # upload/download round trip
    mm = pkg.GpuMmcs(gctx)
    root, tree = mm.commit(dm)
    assert root == ot.root
    for k, layer in enumerate(ot.layers):
        assert tree.layer(k) == layer
    for idx in sorted({0, h - 1, h // 2, (h * 5) // 7}):
        rows, proof = mm.open_batch(idx, tree)
        orows, oproof = ot.open_batch(idx)
        assert rows == orows and proof == oproof
        assert OM.verify_batch(p2params, root, h, idx, rows, proof)
    tree.free()
    for d in dm:
        d.free()


def test_merkle_rejects_mixed_heights(pkg, gctx):
    a = gctx.upload([[1], [2], [3], [4]])
    b = gctx.upload([[1], [2]])
    with pytest.raises(pkg.BackendError):
        pkg.GpuMmcs(gctx).commit([a, b])
