"""BLS12-377 scalar field Fr -- oracle restatement (test infrastructure only).

Reference anchors: `Val = Bls12_377Fr` (`bin/src/config.rs:9`), constructed
from `ark_bls12_377::Fr::from_be_bytes_mod_order` (`trace/src/permutation.rs:102`).
The arithmetic itself is arkworks `Fp256<MontBackend<FrConfig,4>>`
(ark-ff 0.5.0 / ark-bls12-377 0.5.0, `Cargo.lock:51-53,83-85`), which is not
under /root/reference; constants below are checked against the values the
arkworks sources publish (see KAT_* and tests/test_oracle_field.py).

Elements are plain Python ints in [0, R_MOD).  The memory format the CUDA
library and the Rust host share is 4 x u64 little-endian limbs of the
Montgomery representative a*2^256 mod r (`to_mont_limbs`).
"""
from __future__ import annotations

R_MOD = 8444461749428370424248824938781546531375899335154063827935233455917409239041
R_BITS = 253
TWO_ADICITY = 47
GENERATOR = 22  # ark_bls12_377::FrConfig `#[generator = "22"]`
MONT_R = (1 << 256) % R_MOD
MONT_R2 = MONT_R * MONT_R % R_MOD
MONT_RINV = pow(MONT_R, -1, R_MOD)
MONT_NINV64 = (-pow(R_MOD, -1, 1 << 64)) % (1 << 64)
MONT_NINV32 = (-pow(R_MOD, -1, 1 << 32)) % (1 << 32)
TWO_ADIC_ROOT = pow(GENERATOR, (R_MOD - 1) >> TWO_ADICITY, R_MOD)

# Known answers published with the arkworks bls12-377 Fr parameters.
KAT_R_LIMBS = [9015221291577245683, 8239323489949974514, 1646089257421115374, 958099254763297437]
KAT_R2_LIMBS = [2726216793283724667, 14712177743343147295, 12091039717619697043, 81024008013859129]
KAT_NINV64 = 725501752471715839
KAT_GENERATOR_MONT = 5642976643016801619665363617888466827793962762719196659561577942948671127251
KAT_TWO_ADIC_ROOT = 8065159656716812877374967518403273466521432693661810619979959746626482506078

MASK64 = (1 << 64) - 1


def add(a, b):
    return (a + b) % R_MOD


def sub(a, b):
    return (a - b) % R_MOD


def mul(a, b):
    return a * b % R_MOD


def inv(a):
    if a % R_MOD == 0:
        raise ZeroDivisionError("inverse of zero in Fr")
    return pow(a, -1, R_MOD)


def halve(a):
    return a * ((R_MOD + 1) // 2) % R_MOD


class FieldConsts:
    """The two parameters the fork-only `p3-bls12-377-fr` crate fixes and this tree cannot see:
    `Val::GENERATOR` (the coset shift of the PCS) and `two_adic_generator(47)`.  Defaults: arkworks' FrConfig values.
    Mirrors `lsp_set_field_consts`; tests switch them to show that nothing else in the prover depends on the defaults."""
    generator = GENERATOR
    two_adic_root = TWO_ADIC_ROOT


def set_field_consts(generator: int = GENERATOR, two_adic_root: int = TWO_ADIC_ROOT):
    assert generator % R_MOD != 0 and pow(generator, 1 << 31, R_MOD) != 1, "generator lies in a two-adic subgroup"
    assert pow(two_adic_root, 1 << (TWO_ADICITY - 1), R_MOD) == R_MOD - 1, "not a primitive 2^47-th root of unity"
    FieldConsts.generator, FieldConsts.two_adic_root = generator % R_MOD, two_adic_root % R_MOD


def generator() -> int:
    return FieldConsts.generator


def two_adic_generator(bits: int) -> int:
    """omega_{2^bits}: the 2^47-th root squared (47 - bits) times."""
    assert 0 <= bits <= TWO_ADICITY
    return pow(FieldConsts.two_adic_root, 1 << (TWO_ADICITY - bits), R_MOD)


def from_be_bytes_mod_order(b: bytes) -> int:
    """`trace/src/permutation.rs:102`."""
    return int.from_bytes(b, "big") % R_MOD


def to_limbs(x: int) -> list[int]:
    return [(x >> (64 * i)) & MASK64 for i in range(4)]


def from_limbs(l) -> int:
    return int(l[0]) | (int(l[1]) << 64) | (int(l[2]) << 128) | (int(l[3]) << 192)


def to_mont(x: int) -> int:
    return x * MONT_R % R_MOD


def from_mont(x: int) -> int:
    return x * MONT_RINV % R_MOD


def to_mont_limbs(x: int) -> list[int]:
    return to_limbs(to_mont(x % R_MOD))


def from_mont_limbs(l) -> int:
    return from_mont(from_limbs(l))


def reverse_bits_len(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def log2_strict(n: int) -> int:
    assert n > 0 and n & (n - 1) == 0, f"{n} is not a power of two"
    return n.bit_length() - 1


def log2_ceil(n: int) -> int:
    return 0 if n <= 1 else (n - 1).bit_length()


def batch_inverse(xs):
    """Montgomery's trick; all inputs must be non-zero."""
    n = len(xs)
    pref = [1] * (n + 1)
    for i, x in enumerate(xs):
        pref[i + 1] = pref[i] * x % R_MOD
    acc = inv(pref[n])
    out = [0] * n
    for i in range(n - 1, -1, -1):
        out[i] = acc * pref[i] % R_MOD
        acc = acc * xs[i] % R_MOD
    return out


class SplitMix64:
    """Deterministic stand-in for the reference's `thread_rng()` draws
    (`bin/src/main.rs:29-31,49`).  The same generator is restated in
    oracle/c/lsp_oracle.c so Python, C and CUDA runs share constants."""

    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def next_fr(self) -> int:
        """Uniform element of Fr: 253 random bits, rejection-sampled."""
        while True:
            l = [self.next_u64() for _ in range(4)]
            l[3] &= (1 << (R_BITS - 192)) - 1
            v = from_limbs(l)
            if v < R_MOD:
                return v

    def next_below(self, n: int) -> int:
        """Uniform-ish integer in [0, n) (n < 2^63); modulo bias ignored."""
        return self.next_u64() % n
