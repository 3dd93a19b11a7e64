"""Oracle field layer pinned to the constants arkworks publishes for BLS12-377 Fr."""
from oracle import field as F


def test_published_constants():
    assert F.to_limbs(F.MONT_R) == F.KAT_R_LIMBS
    assert F.to_limbs(F.MONT_R2) == F.KAT_R2_LIMBS
    assert F.MONT_NINV64 == F.KAT_NINV64
    assert F.to_mont(F.GENERATOR) == F.KAT_GENERATOR_MONT
    assert F.TWO_ADIC_ROOT == F.KAT_TWO_ADIC_ROOT
    assert F.MONT_NINV32 == 0xFFFFFFFF
    assert F.R_MOD.bit_length() == 253
    assert F.R_MOD == 0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001


def test_two_adicity_and_generator():
    assert (F.R_MOD - 1) % (1 << 47) == 0 and ((F.R_MOD - 1) >> 47) % 2 == 1
    w = F.two_adic_generator(47)
    assert pow(w, 1 << 47, F.R_MOD) == 1 and pow(w, 1 << 46, F.R_MOD) == F.R_MOD - 1
    for k in range(1, 12):
        wk = F.two_adic_generator(k)
        assert pow(wk, 1 << k, F.R_MOD) == 1 and pow(wk, 1 << (k - 1), F.R_MOD) != 1
    # 22 generates the whole multiplicative group: check every prime factor of r-1
    for q in (2, 3, 5, 7, 13, 499, 958612291309063373, 9586122913090633729):
        assert (F.R_MOD - 1) % q == 0
        assert pow(F.GENERATOR, (F.R_MOD - 1) // q, F.R_MOD) != 1


def test_mont_roundtrip_and_bytes():
    rng = F.SplitMix64(3)
    for _ in range(50):
        x = rng.next_fr()
        assert F.from_mont_limbs(F.to_mont_limbs(x)) == x
    assert F.from_be_bytes_mod_order(b"\xff" * 32) == (2**256 - 1) % F.R_MOD
    assert F.from_be_bytes_mod_order((5).to_bytes(32, "big")) == 5


def test_batch_inverse_and_bitrev():
    rng = F.SplitMix64(4)
    xs = [rng.next_fr() or 1 for _ in range(17)]
    assert [x * y % F.R_MOD for x, y in zip(xs, F.batch_inverse(xs))] == [1] * 17
    assert [F.reverse_bits_len(i, 3) for i in range(8)] == [0, 4, 2, 6, 1, 5, 3, 7]
    assert F.halve(7) * 2 % F.R_MOD == 7
