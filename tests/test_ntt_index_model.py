"""A CPU model of `k_ntt_tile`'s index arithmetic (csrc/ntt.cu): how a transform is cut into passes and tiles, which twiddle
of the tile's shared-memory table a butterfly reads, and the radix-4 register blocking (two stages per barrier) -- executed
with the oracle's field arithmetic on sizes where the transform needs THREE passes (a middle pass has non-zero bits both
below and above the tile), and compared with the oracle's DFT.  The GPU tests cover the kernel itself; this pins the
indexing it implements independently of any device."""
import pytest

from oracle import dft as OD
from oracle import field as F

R = F.R_MOD


def split_passes(log_n, max_t):
    k = (log_n + max_t - 1) // max_t
    return [log_n // k + (1 if i < log_n % k else 0) for i in range(k)]


def tile_pass(data, log_n, bit_lo, t, dif, tw):
    """One launch of k_ntt_tile over a single column: every tile of 2^t elements (stride 2^bit_lo)."""
    n, tile = 1 << log_n, 1 << t
    out = list(data)
    for tile_id in range(n >> t):
        idx_lo, idx_hi = tile_id & ((1 << bit_lo) - 1), tile_id >> bit_lo
        base = (idx_hi << (bit_lo + t)) | idx_lo
        # the tile's twiddle table: stage with local bit lb at [2^lb - 1, 2^(lb+1) - 1)
        table = []
        for k in range(tile - 1):
            lb = (k + 1).bit_length() - 1
            u, b = k + 1 - (1 << lb), bit_lo + lb
            table.append(tw[(idx_lo + (u << bit_lo)) << (log_n - 1 - b)])
        x = [out[base + (j << bit_lo)] for j in range(tile)]

        def bfly(i0, i1, w):
            u, v = x[i0], x[i1]
            if dif:
                x[i0], x[i1] = (u + v) % R, (u - v) * w % R
            else:
                vw = v * w % R
                x[i0], x[i1] = (u + vw) % R, (u - vw) % R

        s = 0
        while s + 1 < t:                       # radix-4 rounds: two stages per barrier
            lbA = t - 1 - s if dif else s + 1
            lbB = lbA - 1
            for q in range(tile >> 2):
                low = q & ((1 << lbB) - 1)
                i00 = ((q >> lbB) << (lbB + 2)) | low
                i01, i10 = i00 + (1 << lbB), i00 + (1 << lbA)
                i11 = i10 + (1 << lbB)
                ta, tb = (1 << lbA) - 1, (1 << lbB) - 1
                if dif:
                    bfly(i00, i10, table[ta + low])
                    bfly(i01, i11, table[ta + low + (1 << lbB)])
                    bfly(i00, i01, table[tb + low])
                    bfly(i10, i11, table[tb + low])
                else:
                    bfly(i00, i01, table[tb + low])
                    bfly(i10, i11, table[tb + low])
                    bfly(i00, i10, table[ta + low])
                    bfly(i01, i11, table[ta + low + (1 << lbB)])
            s += 2
        if s < t:                              # odd number of stages: one radix-2 stage left
            lb = 0 if dif else t - 1
            for bf in range(tile >> 1):
                low = bf & ((1 << lb) - 1)
                i0 = ((bf >> lb) << (lb + 1)) | low
                bfly(i0, i0 + (1 << lb), table[(1 << lb) - 1 + low])
        for j in range(tile):
            out[base + (j << bit_lo)] = x[j]
    return out


@pytest.mark.parametrize("log_n,max_t", [(7, 3), (8, 3), (9, 4), (6, 2), (10, 4), (5, 10), (1, 10), (3, 1)])
def test_forward_dif_passes_give_the_bit_reversed_dft(log_n, max_t):
    n = 1 << log_n
    rng = F.SplitMix64(log_n * 31 + max_t)
    coeffs = [rng.next_fr() for _ in range(n)]
    w = F.two_adic_generator(log_n)
    tw = [pow(w, j, R) for j in range(max(1, n // 2))]
    data, bit = coeffs, log_n
    for t in split_passes(log_n, max_t):       # coset_evaluate_blocks: high bits first
        bit -= t
        data = tile_pass(data, log_n, bit, t, True, tw)
    want = OD.ntt(coeffs, w)                   # natural-order evaluations p(w^j)
    assert data == [want[F.reverse_bits_len(i, log_n)] for i in range(n)]


@pytest.mark.parametrize("log_n,max_t", [(7, 3), (8, 3), (9, 4), (6, 2), (4, 10)])
def test_inverse_dit_passes_give_the_coefficients(log_n, max_t):
    n = 1 << log_n
    rng = F.SplitMix64(log_n * 17 + max_t)
    evals = [rng.next_fr() for _ in range(n)]
    w_inv = F.inv(F.two_adic_generator(log_n))
    tw = [pow(w_inv, j, R) for j in range(max(1, n // 2))]
    data = [evals[F.reverse_bits_len(i, log_n)] for i in range(n)]     # bit-reversed gather fused into the first pass' loads
    bit = 0
    for t in split_passes(log_n, max_t):       # interpolate_columns: low bits first
        data = tile_pass(data, log_n, bit, t, False, tw)
        bit += t
    ninv = F.inv(n)
    assert [x * ninv % R for x in data] == OD.idft(evals)
