"""`Radix2DitParallel::coset_lde_batch` -- oracle restatement.

Reference anchors: `Dft = Radix2DitParallel<Val>` (`bin/src/config.rs:22`,
`bin/src/main.rs:52,66`); semantics from published Plonky3 `p3-dft`
(SURVEY.md A.3):  for every column, p = the degree<N interpolant over H_N,
and   out[bitrev_{log L}(j)] = p(shift * omega_L^j),  L = N * 2^added_bits.
`coset_lde_batch_naive` is the O(N*L) definition; `coset_lde_batch` is a
radix-2 restatement checked against it.
"""
from __future__ import annotations

from .field import R_MOD, inv, log2_strict, reverse_bits_len, two_adic_generator


def bit_reverse_rows(rows):
    n = len(rows)
    b = log2_strict(n)
    return [rows[reverse_bits_len(i, b)] for i in range(n)]


def ntt(vals, root):
    """In-order radix-2 DIT: returns [sum_k vals[k] root^(jk)] for j in 0..n."""
    n = len(vals)
    if n == 1:
        return list(vals)
    b = log2_strict(n)
    a = [vals[reverse_bits_len(i, b)] for i in range(n)]
    m = 1
    while m < n:
        wm = pow(root, n // (2 * m), R_MOD)
        for s in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                u = a[s + j]
                t = a[s + j + m] * w % R_MOD
                a[s + j] = (u + t) % R_MOD
                a[s + j + m] = (u - t) % R_MOD
                w = w * wm % R_MOD
        m *= 2
    return a


def idft(vals):
    """Coefficients of the interpolant of `vals` over H_n (natural order)."""
    n = len(vals)
    w = two_adic_generator(log2_strict(n))
    ninv = inv(n % R_MOD)
    return [x * ninv % R_MOD for x in ntt(vals, inv(w))]


def columns_of(mat):
    return [list(c) for c in zip(*mat)] if mat else []


def rows_of(cols):
    return [list(r) for r in zip(*cols)]


def coset_lde_batch(mat, added_bits: int, shift: int):
    """mat: N rows x W.  Returns the L x W matrix in BIT-REVERSED row order
    (the storage the PCS commits to: `.bit_reverse_rows().to_row_major_matrix()`)."""
    n = len(mat)
    log_n = log2_strict(n)
    log_l = log_n + added_bits
    big = 1 << log_l
    w_l = two_adic_generator(log_l)
    out_cols = []
    for col in columns_of(mat):
        coef = idft(col)
        sc = []
        s = 1
        for c in coef:
            sc.append(c * s % R_MOD)
            s = s * shift % R_MOD
        sc += [0] * (big - n)
        ev = ntt(sc, w_l)  # ev[j] = p(shift * w_l^j)
        out_cols.append([ev[reverse_bits_len(i, log_l)] for i in range(big)])
    return rows_of(out_cols)


def coset_lde_batch_naive(mat, added_bits: int, shift: int):
    """Definition: Lagrange interpolation over H_N, pointwise evaluation."""
    n = len(mat)
    log_n = log2_strict(n)
    log_l = log_n + added_bits
    big = 1 << log_l
    w_n = two_adic_generator(log_n)
    w_l = two_adic_generator(log_l)
    ninv = inv(n % R_MOD)
    out_cols = []
    for col in columns_of(mat):
        coef = []
        for k in range(n):  # coef_k = 1/n sum_i col[i] w^{-ik}
            acc = 0
            wk = pow(w_n, (-k) % n, R_MOD)
            x = 1
            for i in range(n):
                acc = (acc + col[i] * x) % R_MOD
                x = x * wk % R_MOD
            coef.append(acc * ninv % R_MOD)
        ev = []
        for j in range(big):
            x = shift * pow(w_l, j, R_MOD) % R_MOD
            acc = 0
            for c in reversed(coef):
                acc = (acc * x + c) % R_MOD
            ev.append(acc)
        out_cols.append([ev[reverse_bits_len(i, log_l)] for i in range(big)])
    return rows_of(out_cols)


def eval_poly(coef, x):
    acc = 0
    for c in reversed(coef):
        acc = (acc * x + c) % R_MOD
    return acc
