"""GPU parity tests of the batch codecs against the CPU oracle:
CompressedEdwardsY::decompress (C/edwards.rs:211-257), EdwardsPoint::compress_batch (C/edwards.rs:619-647),
CompressedRistretto::decompress (C/ristretto.rs:266-345), RistrettoPoint::double_and_compress_batch
(C/ristretto.rs:564-646; reference test double_and_compress_1024_random_points, C/ristretto.rs:1683-1698)."""
import ctypes as C
import random

import pytest

import pyref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import curve25519_dalek_b200 as pkg
    e = pkg.Engine(0)
    yield e
    e.close()


def b32(x):
    return x.to_bytes(32, "little")


def limbs_of(oracle, pts):
    buf = (C.c_uint64 * (20 * max(len(pts), 1)))()
    for i, p in enumerate(pts):
        for k, v in enumerate(oracle.p3_limbs(p)):
            buf[20 * i + k] = v
    return buf


@pytest.mark.parametrize("f64", [1, 0])
def test_edwards_decompress_batch(eng, oracle, f64):
    p = pyref.p
    rnd = random.Random(5)
    encs = [b32(y) for y in (0, 1, 2, 3, 4, p - 1, p, p + 1, 2**255 - 1, 2**255 - 20)]
    encs += [b32(y | (1 << 255)) for y in (0, 1, p - 1, p, 4)]
    encs += [oracle.compress(oracle.scalarmul(b32(rnd.randrange(pyref.L)), oracle.basepoint())) for _ in range(40)]
    encs += [rnd.randbytes(32) for _ in range(80)]
    n = len(encs)
    eng.set_option("decompress_f64", f64)
    try:
        rc, limbs, ok = eng.decompress_batch(b"".join(encs), n)
    finally:
        eng.set_option("decompress_f64", 1)
    want = [oracle.decompress(e) for e in encs]
    assert [bool(x) for x in ok] == [w is not None for w in want]
    assert rc == (0 if all(ok) else 1)
    for i, w in enumerate(want):
        got = oracle.p3_from_limbs(list(limbs[20 * i:20 * i + 20]))
        assert oracle.compress(got) == (oracle.compress(w) if w is not None else oracle.compress(oracle.identity()))
    assert eng.decompress_batch(b"", 0)[0] == 0


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 500, 1031])
def test_edwards_compress_batch(eng, oracle, n):
    """compress_batch == compress of every point, for Z != 1 representatives, identity and small-order points,
    across the 8-point inversion groups (ragged tail included)."""
    rnd = random.Random(n)
    B = oracle.basepoint()
    pts = []
    for i in range(n):
        q = oracle.scalarmul(b32(rnd.randrange(pyref.L)), B)
        pts.append(oracle.sub(oracle.add(oracle.double(q), q), oracle.double(q)))       # same point, Z != 1
    if n >= 9:
        pts[0] = oracle.identity(); pts[3] = oracle.decompress(b32(0)); pts[8] = oracle.decompress(b32(pyref.p - 1))
    got = eng.compress_batch(limbs_of(oracle, pts), n)
    assert got == oracle.compress_batch(pts) == b"".join(oracle.compress(q) for q in pts)


def test_ristretto_decompress_and_double_compress_batch(eng, oracle, kat):
    rnd = random.Random(9)
    encs = [bytes.fromhex(h) for h in kat["ristretto"]["SMALL_MULTIPLES"]["hex"]]
    G = oracle.ristretto_decompress(encs[1])
    encs += [oracle.ristretto_compress(oracle.scalarmul(b32(rnd.randrange(pyref.L)), G)) for _ in range(100)]
    encs += [rnd.randbytes(32) for _ in range(60)] + [b32(1), b32(pyref.p - 1), b32(pyref.p), b32(2**255 - 1)]
    n = len(encs)
    rc, limbs, ok = eng.decompress_batch(b"".join(encs), n, ristretto=True)
    want = [oracle.ristretto_decompress(e) for e in encs]
    assert [bool(x) for x in ok] == [w is not None for w in want] and rc == 1
    good = [i for i in range(n) if ok[i]]
    for i in good:                                             # same coset: re-encoding gives the input back
        assert oracle.ristretto_compress(oracle.p3_from_limbs(list(limbs[20 * i:20 * i + 20]))) == encs[i]
    # double_and_compress over the decoded points (identity included: encs[0]) and over Z != 1 representatives
    pts = [want[i] for i in good]
    pts += [oracle.sub(oracle.add(oracle.double(q), q), oracle.double(q)) for q in pts[:50]]
    got = eng.ristretto_double_and_compress_batch(limbs_of(oracle, pts), len(pts))
    assert got == oracle.ristretto_double_and_compress_batch(pts) == b"".join(oracle.ristretto_compress(oracle.double(q)) for q in pts)
    # doubling the small multiples i*G gives the encodings of 2i*G (C/ristretto.rs:1387-1461)
    small = [oracle.ristretto_decompress(e) for e in encs[:8]]
    got = eng.ristretto_double_and_compress_batch(limbs_of(oracle, small), 8)
    assert [got[32 * i:32 * i + 32] for i in range(8)] == [encs[2 * i] for i in range(8)]


def test_codecs_large_round_trip(eng, oracle):
    """2^18 points made on the GPU: compress_batch then decompress_batch returns the same points (compared through a
    second compression), spot-checked against the oracle; pieces of 2^16 are streamed over two streams."""
    import numpy as np
    n = (1 << 18) + 5
    rng = np.random.Generator(np.random.PCG64(3))
    t = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); t[:, 31] &= 0x0F
    limbs, comp = eng.mul_base_batch(t, n)
    enc = eng.compress_batch(limbs, n)
    assert enc == comp                                         # the fixed-base kernel's own encodings
    rc, limbs2, ok = eng.decompress_batch(enc, n)
    assert rc == 0 and all(ok)
    assert eng.compress_batch(limbs2, n) == enc
    for i in (0, 65535, 65536, 131072, n - 1):
        assert enc[32 * i:32 * i + 32] == oracle.compress(oracle.scalarmul(t[i].tobytes(), oracle.basepoint()))
    dbl = eng.ristretto_double_and_compress_batch(limbs, n)
    for i in (0, 65535, 65536, n - 1):
        q = oracle.scalarmul(t[i].tobytes(), oracle.basepoint())
        assert dbl[32 * i:32 * i + 32] == oracle.ristretto_compress(oracle.double(q))


def test_scalar_batches(eng, oracle):
    """Scalar::from_bytes_mod_order_wide and Scalar::invert_batch against the oracle / big integers."""
    import curve25519_dalek_b200 as pkg
    rnd = random.Random(12)
    wide = [rnd.randbytes(64) for _ in range(300)] + [bytes(64), b"\xff" * 64, pyref.L.to_bytes(64, "little")]
    got = eng.scalar_from_wide_batch(b"".join(wide), len(wide))
    assert [int.from_bytes(got[32 * i:32 * i + 32], "little") for i in range(len(wide))] == [int.from_bytes(w, "little") % pyref.L for w in wide]
    for n in (0, 1, 7, 8, 9, 1000):
        xs = [rnd.randrange(1, 2**255) for _ in range(n)]
        xs = [x if x % pyref.L else 1 for x in xs]
        if n >= 9:
            xs[0], xs[1], xs[8] = 1, pyref.L - 1, pyref.L + 5
        inv, prod = eng.scalar_invert_batch(b"".join(b32(x) for x in xs), n)
        want_inv, want_prod = oracle.scalar_invert_batch([b32(x) for x in xs])
        assert inv == b"".join(want_inv) and prod == want_prod
        assert all(int.from_bytes(inv[32 * i:32 * i + 32], "little") * xs[i] % pyref.L == 1 for i in range(n))
    with pytest.raises(pkg.EngineError):                      # a zero input violates the precondition
        eng.scalar_invert_batch(b"".join(b32(x) for x in (5, pyref.L, 7)), 3)
