"""Pin the CPU oracle against every known-answer vector the reference holds for the hot path
 and against an independent big-integer model.  CPU only."""
import hashlib
import json
import os
import random

import pytest

import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = bytes.fromhex


def g(kat, group, name):
    return H(kat[group][name]["hex"])


# ---------------------------------------------------------------- constants
def test_constants_match_reference_limbs(kat):
    """oracle/constants.h (derived by big-int arithmetic) == u64/constants.rs limb tables."""
    import re
    txt = open(os.path.join(ROOT, "oracle", "constants.h")).read()

    def limbs(name):
        m = re.search(r"%s\[5\] = \{([^}]*)\}" % name, txt)
        return [int(x) for x in re.findall(r"UINT64_C\((\d+)\)", m.group(1))]
    c = kat["u64_constants"]
    assert limbs("K_EDWARDS_D") == c["EDWARDS_D"]["limbs"]
    assert limbs("K_EDWARDS_D2") == c["EDWARDS_D2"]["limbs"]
    assert limbs("K_SQRT_M1") == c["SQRT_M1"]["limbs"]
    assert limbs("K_INVSQRT_A_MINUS_D") == c["INVSQRT_A_MINUS_D"]["limbs"]
    assert limbs("K_SC_L") == c["L"]["limbs"]
    assert limbs("K_SC_R") == c["R"]["limbs"]
    assert limbs("K_SC_RR") == c["RR"]["limbs"]
    m = re.search(r"K_SC_LFACTOR = UINT64_C\((\d+)\)", txt)
    assert int(m.group(1)) == c["LFACTOR"]["limbs"][0]
    bp = c["ED25519_BASEPOINT_POINT_XYZT"]["limbs"]
    assert limbs("K_BASE_X") == bp[0:5] and limbs("K_BASE_Y") == bp[5:10]
    assert bp[10:15] == [1, 0, 0, 0, 0] and limbs("K_BASE_T") == bp[15:20]


# ---------------------------------------------------------------- field (C/field.rs:522-698)
def test_field_kats(oracle, kat):
    a = oracle.fe_from_bytes(g(kat, "field", "A_BYTES"))
    assert oracle.fe_to_bytes(oracle.fe_op2("fe_mul", a, a)) == g(kat, "field", "ASQ_BYTES")
    assert oracle.fe_to_bytes(oracle.fe_op1("fe_square", a)) == g(kat, "field", "ASQ_BYTES")
    asq = oracle.fe_from_bytes(g(kat, "field", "ASQ_BYTES"))
    sq2 = oracle.fe_op1("fe_square2", a)
    assert oracle.fe_to_bytes(sq2) == oracle.fe_to_bytes(oracle.fe_op2("fe_add", asq, asq))
    assert oracle.fe_to_bytes(oracle.fe_op1("fe_invert", a)) == g(kat, "field", "AINV_BYTES")
    assert oracle.fe_to_bytes(oracle.fe_op1("fe_pow_p58", a)) == g(kat, "field", "AP58_BYTES")
    ainv = oracle.fe_from_bytes(g(kat, "field", "AINV_BYTES"))
    one = oracle.fe_op2("fe_mul", a, ainv)
    assert oracle.fe_to_bytes(one) == (1).to_bytes(32, "little")


def test_field_highbit_ignored_and_noncanonical(oracle, kat):
    b = bytearray(g(kat, "field", "B_BYTES"))
    with_hi = oracle.fe_to_bytes(oracle.fe_from_bytes(bytes(b)))
    b[31] &= 127
    assert with_hi == oracle.fe_to_bytes(oracle.fe_from_bytes(bytes(b)))
    # 2^255 - 18 decodes to 1 (C/field.rs:682-698)
    enc = (2**255 - 18).to_bytes(32, "little")
    assert oracle.fe_to_bytes(oracle.fe_from_bytes(enc)) == (1).to_bytes(32, "little")


def test_sqrt_ratio_behaviour(oracle):
    """C/field.rs:599-636"""
    def fe(x):
        return oracle.fe_from_bytes((x % pyref.p).to_bytes(32, "little"))

    def val(f):
        return int.from_bytes(oracle.fe_to_bytes(f), "little")
    ok, r = oracle.fe_sqrt_ratio_i(fe(0), fe(0)); assert ok and val(r) == 0
    ok, r = oracle.fe_sqrt_ratio_i(fe(1), fe(0)); assert not ok and val(r) == 0
    ok, r = oracle.fe_sqrt_ratio_i(fe(2), fe(1))
    assert not ok and val(r) ** 2 % pyref.p == 2 * pyref.SQRT_M1 % pyref.p and val(r) % 2 == 0
    ok, r = oracle.fe_sqrt_ratio_i(fe(4), fe(1)); assert ok and val(r) ** 2 % pyref.p == 4 and val(r) % 2 == 0
    ok, r = oracle.fe_sqrt_ratio_i(fe(1), fe(4)); assert ok and val(r) ** 2 * 4 % pyref.p == 1 and val(r) % 2 == 0


def test_field_random_vs_bigint(oracle):
    rnd = random.Random(1)
    for _ in range(300):
        x, y = rnd.randrange(2**255), rnd.randrange(2**255)
        fx, fy = oracle.fe_from_bytes(x.to_bytes(32, "little")), oracle.fe_from_bytes(y.to_bytes(32, "little"))
        for name, f in (("fe_mul", lambda a, b: a * b), ("fe_add", lambda a, b: a + b), ("fe_sub", lambda a, b: a - b)):
            got = int.from_bytes(oracle.fe_to_bytes(oracle.fe_op2(name, fx, fy)), "little")
            assert got == f(x, y) % pyref.p
        assert int.from_bytes(oracle.fe_to_bytes(oracle.fe_op1("fe_neg", fx)), "little") == (-x) % pyref.p


# ---------------------------------------------------------------- scalars (C/scalar.rs:1425-1963)
def test_scalar_kats(oracle, kat):
    X, Y = g(kat, "scalar", "X"), g(kat, "scalar", "Y")
    assert oracle.sc_op2("scalar_mul", X, Y) == g(kat, "scalar", "X_TIMES_Y")
    assert oracle.sc_op1("scalar_invert", X) == g(kat, "scalar", "XINV")
    assert oracle.sc_op1("scalar_reduce", b"\xff" * 32) == g(kat, "scalar", "CANONICAL_2_256_MINUS_1")
    assert oracle.scalar_from_wide(X + X) == g(kat, "scalar", "X_PLUS_2_256_X_REDUCED")
    assert oracle.naf(g(kat, "scalar", "A_SCALAR"), 5) == kat["scalar"]["A_NAF"]["ints"]
    # neg / add / sub vs bigint
    x, y = pyref.sc(X), pyref.sc(Y)
    assert pyref.sc(oracle.sc_op2("scalar_add", X, Y)) == (x + y) % pyref.L
    assert pyref.sc(oracle.sc_op2("scalar_sub", X, Y)) == (x - y) % pyref.L
    assert pyref.sc(oracle.sc_op1("scalar_neg", X)) == (-x) % pyref.L
    assert oracle.scalar_is_canonical(X)
    assert not oracle.scalar_is_canonical(pyref.L.to_bytes(32, "little"))
    assert oracle.scalar_is_canonical((pyref.L - 1).to_bytes(32, "little"))
    assert not oracle.scalar_is_canonical(g(kat, "scalar", "LARGEST_UNREDUCED_SCALAR"))


def test_scalar_wide_reduce_random(oracle):
    rnd = random.Random(2)
    for _ in range(200):
        b = rnd.randbytes(64)
        assert pyref.sc(oracle.scalar_from_wide(b)) == int.from_bytes(b, "little") % pyref.L


@pytest.mark.parametrize("w", [5, 6, 7, 8])
def test_naf_roundtrip(oracle, w):
    """C/scalar.rs:1561-1589"""
    rnd = random.Random(w)
    for _ in range(200):
        x = rnd.randrange(pyref.L)
        naf = oracle.naf(x.to_bytes(32, "little"), w)
        assert sum(dg << i for i, dg in enumerate(naf)) == x
        lim = 1 << (w - 1)
        for i, dg in enumerate(naf):
            assert dg == 0 or (dg % 2 == 1 and -lim < dg < lim)


@pytest.mark.parametrize("w", [4, 5, 6, 7, 8])
def test_radix_2w_roundtrip(oracle, kat, w):
    """C/scalar.rs:1923-1963 incl. the largest unreduced scalar"""
    rnd = random.Random(100 + w)
    cases = [rnd.randrange(2**255) for _ in range(100)] + [0, 1, pyref.L - 1, 2**255 - 1]
    hint = oracle.radix_size_hint(w)
    for x in cases:
        digits = oracle.radix2w(x.to_bytes(32, "little"), w)
        assert sum(dg << (w * i) for i, dg in enumerate(digits)) == x
        assert all(dg == 0 for dg in digits[hint:])
        lim = 1 << (w - 1)
        assert all(-lim <= dg < lim for dg in digits[:hint - 1]) and -lim <= digits[hint - 1] <= lim


# ---------------------------------------------------------------- edwards (C/edwards.rs:1806-2451)
def test_basepoint_and_small_multiples(oracle, kat):
    B = oracle.basepoint()
    assert oracle.compress(B) == g(kat, "constants", "ED25519_BASEPOINT_COMPRESSED")
    bx = oracle.fe_to_bytes(B.X)
    assert bx == g(kat, "edwards", "BASE_X_COORD_BYTES")
    d = oracle.decompress(g(kat, "constants", "ED25519_BASEPOINT_COMPRESSED"))
    assert oracle.p3_limbs(d) == oracle.p3_limbs(B)    # limb-exact, as the reference constant
    assert oracle.compress(oracle.double(B)) == g(kat, "edwards", "BASE2_CMPRSSD")
    assert oracle.compress(oracle.add(B, B)) == g(kat, "edwards", "BASE2_CMPRSSD")
    assert oracle.compress(oracle.mul_by_pow_2(B, 4)) == g(kat, "edwards", "BASE16_CMPRSSD")
    # sign bit round trip (C/edwards.rs:1861-1887)
    negB = oracle.decompress(bytes(oracle.compress(B)[:31]) + bytes([oracle.compress(B)[31] | 0x80]))
    assert oracle.is_identity(oracle.add(B, negB))
    assert oracle.compress(oracle.identity()) == (1).to_bytes(32, "little")


def test_scalar_mul_and_double_scalar_kats(oracle, kat):
    B = oracle.basepoint()
    a, b = g(kat, "edwards", "A_SCALAR"), g(kat, "edwards", "B_SCALAR")
    aB = oracle.scalarmul(a, B)
    assert oracle.compress(aB) == g(kat, "edwards", "A_TIMES_BASEPOINT")
    A = oracle.decompress(g(kat, "edwards", "A_TIMES_BASEPOINT"))
    want = g(kat, "edwards", "DOUBLE_SCALAR_MULT_RESULT")
    # multiscalar_mul_vs_ed25519py (:2428-2435) through every algorithm
    assert oracle.compress(oracle.msm("optional", [a, b], [A, B])) == want
    assert oracle.compress(oracle.msm("straus_vartime", [a, b], [A, B])) == want
    assert oracle.compress(oracle.msm("pippenger", [a, b], [A, B])) == want
    assert oracle.compress(oracle.msm_ct([a, b], [A, B])) == want      # vartime_vs_consttime (:2439-2451)


def test_pippenger_reference_test(oracle):
    """scalar_mul/pippenger.rs:169-198: P_i=(1+i)B, s_i = 1/2128506 + i/4443282, n=512..1."""
    B = oracle.basepoint()
    x = oracle.sc_op1("scalar_invert", (2128506).to_bytes(32, "little"))
    y = oracle.sc_op1("scalar_invert", (4443282).to_bytes(32, "little"))
    n = 512
    points, scalars, acc = [], [], B
    for i in range(n):
        points.append(acc); acc = oracle.add(acc, B)
        scalars.append(oracle.sc_op2("scalar_add", x, oracle.sc_op2("scalar_mul", (i).to_bytes(32, "little"), y)))
    # control via big-int scalars: sum s_i * (1+i) * B
    while n > 0:
        k = sum(pyref.sc(scalars[i]) * (1 + i) for i in range(n)) % pyref.L
        control = oracle.compress(oracle.scalarmul(k.to_bytes(32, "little"), B))
        assert oracle.compress(oracle.msm("pippenger", scalars[:n], points[:n])) == control
        n //= 2


@pytest.mark.parametrize("n", [0, 1, 2, 3, 100, 189, 190, 250, 500, 800, 1000])
def test_msm_algebraic_identity(oracle, n):
    """C/edwards.rs:2276-2335: G_i = x_i B, check sum x_i G_i == (sum x_i^2) B for vartime and
    const-time paths; sizes cross the Straus/Pippenger and w=6/7/8 thresholds."""
    rnd = random.Random(n)
    B = oracle.basepoint()
    xs = [rnd.randrange(pyref.L) for _ in range(n)]
    Gs = [oracle.scalarmul(x.to_bytes(32, "little"), B) for x in xs]
    want = oracle.compress(oracle.scalarmul((sum(x * x for x in xs) % pyref.L).to_bytes(32, "little"), B))
    sc = [x.to_bytes(32, "little") for x in xs]
    assert oracle.compress(oracle.msm("optional", sc, Gs)) == want
    if n <= 250:
        assert oracle.compress(oracle.msm_ct(sc, Gs)) == want
    if n:
        assert oracle.msm("optional", sc, Gs[:-1] + [None]) is None      # any None -> None


def test_points_vs_bigint_model(oracle):
    rnd = random.Random(7)
    B = oracle.basepoint()
    for _ in range(8):
        s, t = rnd.randrange(2**255), rnd.randrange(pyref.L)
        P = oracle.scalarmul(t.to_bytes(32, "little"), B)
        got = oracle.compress(oracle.scalarmul(s.to_bytes(32, "little"), P))
        assert got == pyref.compress(pyref.mul(s * t % pyref.L, pyref.B))
        assert oracle.compress(oracle.decompress(got)) == got


def test_eight_torsion_and_noncanonical_points(oracle, kat):
    limbs = kat["u64_constants"]["EIGHT_TORSION_XYZT"]["limbs"]
    pts = [oracle.p3_from_limbs(limbs[20 * i:20 * i + 20]) for i in range(8)]
    assert oracle.is_identity(pts[0])
    for i, P in enumerate(pts):
        assert oracle.is_identity(oracle.mul_by_pow_2(P, 3))
        enc = oracle.compress(P)
        assert oracle.compress(oracle.decompress(enc)) == enc
    # non-canonical y = p + 1 (== 1, identity) is accepted by decompress (C/edwards.rs:211-257)
    assert oracle.is_identity(oracle.decompress((pyref.p + 1).to_bytes(32, "little")))
    # y = 2 is not on the curve
    assert oracle.decompress((2).to_bytes(32, "little")) is None


# ---------------------------------------------------------------- ristretto (C/ristretto.rs:1351-1483)
def test_ristretto_small_multiples(oracle, kat):
    encs = [H(h) for h in kat["ristretto"]["SMALL_MULTIPLES"]["hex"]]
    Bc = g(kat, "constants", "RISTRETTO_BASEPOINT_COMPRESSED")
    assert encs[1] == Bc
    B = oracle.ristretto_decompress(Bc)
    P = oracle.identity()
    for i in range(16):
        assert oracle.ristretto_compress(P) == encs[i]
        Q = oracle.ristretto_decompress(encs[i])
        assert Q is not None and oracle.ristretto_ct_eq(P, Q)
        P = oracle.add(P, B)
    # the Ristretto basepoint is the Ed25519 basepoint (C/constants.rs:66)
    assert oracle.ristretto_ct_eq(B, oracle.basepoint())


def test_ristretto_torsion_invariance_and_bad_encodings(oracle, kat):
    limbs = kat["u64_constants"]["EIGHT_TORSION_XYZT"]["limbs"]
    rnd = random.Random(3)
    P = oracle.scalarmul(rnd.randrange(pyref.L).to_bytes(32, "little"), oracle.basepoint())
    enc = oracle.ristretto_compress(P)
    for i in (0, 2, 4, 6):   # 4-torsion coset (C/ristretto.rs:1464-1483)
        T4 = oracle.p3_from_limbs(limbs[20 * i:20 * i + 20])
        assert oracle.ristretto_compress(oracle.add(P, T4)) == enc
    # negative s, non-canonical s are rejected
    assert oracle.ristretto_decompress((1).to_bytes(32, "little")) is None
    assert oracle.ristretto_decompress((pyref.p).to_bytes(32, "little")) is None
    assert oracle.ristretto_decompress(b"\xff" * 32) is None


def test_ristretto_double_base_batch(oracle, kat):
    rnd = random.Random(11)
    G = g(kat, "constants", "RISTRETTO_BASEPOINT_COMPRESSED")
    Gp = oracle.ristretto_decompress(G)
    h = rnd.randrange(pyref.L)
    Hp = oracle.scalarmul(h.to_bytes(32, "little"), Gp)
    Hc = oracle.ristretto_compress(Hp)
    n = 6
    a = [rnd.randrange(pyref.L) for _ in range(n)]
    b = [rnd.randrange(pyref.L) for _ in range(n)]
    rc, out = oracle.ristretto_double_base_batch(b"".join(x.to_bytes(32, "little") for x in a),
                                                 b"".join(x.to_bytes(32, "little") for x in b), G, Hc)
    assert rc == 0
    for i in range(n):
        want = oracle.ristretto_compress(oracle.scalarmul(((a[i] + b[i] * h) % pyref.L).to_bytes(32, "little"), Gp))
        assert out[32 * i:32 * i + 32] == want


# ---------------------------------------------------------------- hashing
def test_sha512_vs_hashlib(oracle):
    rnd = random.Random(5)
    for n in [0, 1, 55, 111, 112, 113, 123, 127, 128, 129, 239, 240, 241, 1000]:
        m = rnd.randbytes(n)
        assert oracle.sha512(m) == hashlib.sha512(m).digest()


def test_keccak_f1600_vs_hashlib_sha3(oracle):
    """SHA3-256 built on the oracle's permutation must equal hashlib's."""
    def sha3_256(msg):
        rate = 136
        m = bytearray(msg) + b"\x06"
        while len(m) % rate:
            m += b"\x00"
        m[-1] |= 0x80
        st = [0] * 25
        for off in range(0, len(m), rate):
            for i in range(rate // 8):
                st[i] ^= int.from_bytes(m[off + 8 * i:off + 8 * i + 8], "little")
            st = oracle.keccak_f1600(st)
        return b"".join(x.to_bytes(8, "little") for x in st)[:32]
    rnd = random.Random(6)
    for n in [0, 1, 135, 136, 137, 500]:
        m = rnd.randbytes(n)
        assert sha3_256(m) == hashlib.sha3_256(m).digest()


def test_precomputed_straus_and_batch_codecs(oracle, kat):
    """The oracle's restatements of the SURVEY 8f rows against properties the reference tests:
    precomputed == non-precomputed MSM (C/edwards.rs:2343-2417), compress_batch == compress of each point
    (C/edwards.rs:1885-1901), double_and_compress_batch == compress(2P) (C/ristretto.rs:1683-1698)."""
    rnd = random.Random(2026)
    b32 = lambda x: x.to_bytes(32, "little")
    B = oracle.basepoint()
    sp = [oracle.scalarmul(b32(rnd.randrange(pyref.L)), B) for _ in range(9)]
    dp = [oracle.scalarmul(b32(rnd.randrange(pyref.L)), B) for _ in range(5)]
    sp[2] = oracle.identity(); dp[1] = oracle.decompress(b32(0))
    ss = [b32(rnd.randrange(pyref.L)) for _ in range(9)]; ss[0] = b32(0); ss[1] = b32(2**255 - 1)
    ds = [b32(rnd.randrange(pyref.L)) for _ in range(5)]
    want = oracle.compress(oracle.msm("straus_vartime", ss + ds, sp + dp))
    assert oracle.compress(oracle.precomputed_straus(ss, sp, ds, dp)) == want
    # fewer static scalars than points: the tail is ignored; more: error; a None dynamic point: None
    assert oracle.compress(oracle.precomputed_straus(ss[:4], sp, ds, dp)) == oracle.compress(oracle.msm("straus_vartime", ss[:4] + ds, sp[:4] + dp))
    with pytest.raises(ValueError):
        oracle.precomputed_straus(ss + [b32(1)], sp, ds, dp)
    assert oracle.precomputed_straus(ss, sp, ds, dp[:2] + [None] + dp[3:]) is None
    assert oracle.compress(oracle.precomputed_straus([], [], [], [])) == oracle.compress(oracle.identity())
    # the reference's KAT: A_SCALAR * A_TIMES_BASEPOINT-style double-scalar result with B static, A dynamic
    H = bytes.fromhex
    a, b = H(kat["edwards"]["A_SCALAR"]["hex"]), H(kat["edwards"]["B_SCALAR"]["hex"])
    A = oracle.decompress(H(kat["edwards"]["A_TIMES_BASEPOINT"]["hex"]))
    assert oracle.compress(oracle.precomputed_straus([b], [B], [a], [A])) == H(kat["edwards"]["DOUBLE_SCALAR_MULT_RESULT"]["hex"])
    # codecs
    pts = [oracle.sub(oracle.add(oracle.double(q), q), oracle.double(q)) for q in sp + dp] + [oracle.identity()]
    assert oracle.compress_batch(pts) == b"".join(oracle.compress(q) for q in pts)
    assert oracle.compress_batch([]) == b""
    assert oracle.ristretto_double_and_compress_batch(pts) == b"".join(oracle.ristretto_compress(oracle.double(q)) for q in pts)
    encs = [H(h) for h in kat["ristretto"]["SMALL_MULTIPLES"]["hex"]]
    small = [oracle.ristretto_decompress(e) for e in encs[:8]]
    got = oracle.ristretto_double_and_compress_batch(small)
    assert [got[32 * i:32 * i + 32] for i in range(8)] == [encs[2 * i] for i in range(8)]


def test_scalar_invert_batch(oracle):
    """Scalar::invert_batch against the doc example of C/scalar.rs:765-778 (3, 5, 7, 11) and big-integer inverses."""
    b32 = lambda x: x.to_bytes(32, "little")
    xs = [3, 5, 7, 11, pyref.L - 1, 2**255 - 1, 1]
    inv, prod = oracle.scalar_invert_batch([b32(x) for x in xs])
    assert [int.from_bytes(v, "little") for v in inv] == [pow(x % pyref.L, -1, pyref.L) for x in xs]
    want = 1
    for x in xs:
        want = want * pow(x % pyref.L, -1, pyref.L) % pyref.L
    assert int.from_bytes(prod, "little") == want
    assert oracle.scalar_invert_batch([])[1] == b32(1)
