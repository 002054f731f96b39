"""GPU parity tests: constant-time MSM, Ristretto double-base batch / vartime MSM (config 5),
fixed-base multiples and the GPU signer (input synthesis) against the CPU oracle."""
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


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 100, 250])
def test_ct_msm_matches_oracle(eng, oracle, n):
    """multiscalar_mul (C/edwards.rs:970-995) incl. multiscalar_mul_vartime_vs_consttime (:2439-2451)."""
    rnd = random.Random(n)
    B = oracle.basepoint()
    scalars = [b32(rnd.randrange(pyref.L)) for _ in range(n)]
    if n >= 3:
        scalars[0] = b32(0); scalars[1] = b32(2**255 - 1); scalars[2] = b32(pyref.L - 1)
    pts = [oracle.scalarmul(b32(rnd.randrange(pyref.L)), B) for _ in range(n)]
    if n >= 17:
        pts[5] = oracle.identity(); pts[6] = oracle.decompress(b32(0))
    want = oracle.compress(oracle.msm_ct(scalars, pts))
    rc, got, limbs = eng.edwards_ct_msm(b"".join(scalars), b"".join(oracle.compress(p) for p in pts), n, want_limbs=True)
    assert rc == 0 and got == want
    assert oracle.compress(oracle.p3_from_limbs(limbs)) == want
    ext = (C.c_uint64 * (20 * max(n, 1)))()
    for i, p in enumerate(pts):
        q = oracle.sub(oracle.add(oracle.double(p), p), oracle.double(p))
        for k, v in enumerate(oracle.p3_limbs(q)):
            ext[20 * i + k] = v
    rc, got2, _ = eng.edwards_ct_msm(b"".join(scalars), ext, n, point_fmt=1)
    assert rc == 0 and got2 == want
    # the vartime engine path agrees
    rc, got3, _ = eng.edwards_vartime_msm(b"".join(scalars), b"".join(oracle.compress(p) for p in pts), n)
    assert got3 == want


def test_ct_msm_reference_kat(eng, kat):
    H = bytes.fromhex
    a, b = H(kat["edwards"]["A_SCALAR"]["hex"]), H(kat["edwards"]["B_SCALAR"]["hex"])
    A = H(kat["edwards"]["A_TIMES_BASEPOINT"]["hex"])
    Bc = H(kat["constants"]["ED25519_BASEPOINT_COMPRESSED"]["hex"])
    rc, got, _ = eng.edwards_ct_msm(a + b, A + Bc, 2)
    assert rc == 0 and got == H(kat["edwards"]["DOUBLE_SCALAR_MULT_RESULT"]["hex"])


def test_ristretto_double_base_batch(eng, oracle, kat):
    """config 5: RistrettoPoint::multiscalar_mul([a_i, b_i], [G, H]).compress() per pair."""
    rnd = random.Random(55)
    G = bytes.fromhex(kat["constants"]["RISTRETTO_BASEPOINT_COMPRESSED"]["hex"])
    Gp = oracle.ristretto_decompress(G)
    Hc = oracle.ristretto_compress(oracle.scalarmul(b32(pyref.labelled_scalar(b"dalek-b200/H", 1, 0)), Gp))
    n = 300
    a = [rnd.randrange(pyref.L) for _ in range(n)]
    b = [rnd.randrange(pyref.L) for _ in range(n)]
    a[0], b[0] = 0, 0
    a[1], b[1] = 1, 0
    a[2], b[2] = pyref.L - 1, 2**255 - 1
    ab, bb = b"".join(b32(x) for x in a), b"".join(b32(x) for x in b)
    rc_o, want = oracle.ristretto_double_base_batch(ab, bb, G, Hc)
    rc, got = eng.ristretto_double_base_batch(ab, bb, G, Hc, n)
    assert rc == rc_o == 0 and got == want
    assert got[:32] == bytes(32)                         # identity encodes as zeros
    # invalid base -> None
    rc, _ = eng.ristretto_double_base_batch(ab, bb, G, b32(1), n)
    assert rc == 1
    # encodings of small multiples (C/ristretto.rs:1387-1461) through the engine: i*G + 0*H
    encs = [bytes.fromhex(h) for h in kat["ristretto"]["SMALL_MULTIPLES"]["hex"]]
    rc, got = eng.ristretto_double_base_batch(b"".join(b32(i) for i in range(16)), bytes(32 * 16), G, Hc, 16)
    assert [got[32 * i:32 * i + 32] for i in range(16)] == encs


@pytest.mark.parametrize("variant", [0, 1])
def test_ristretto_double_base_comb(eng, oracle, kat, variant):
    """Batches of >= 4096 pairs go through the shared-memory fixed-base comb; every variant must give
    the oracle's bytes (Straus restatement of straus.rs:106-146), edge scalars included."""
    rnd = random.Random(57)
    G = bytes.fromhex(kat["constants"]["RISTRETTO_BASEPOINT_COMPRESSED"]["hex"])
    Gp = oracle.ristretto_decompress(G)
    Hc = oracle.ristretto_compress(oracle.scalarmul(b32(pyref.labelled_scalar(b"dalek-b200/H", 1, 0)), Gp))
    n = 4200
    a = [rnd.randrange(pyref.L) for _ in range(n)]
    b = [rnd.randrange(pyref.L) for _ in range(n)]
    edge = [0, 1, 8, 9, 15, 16, pyref.L - 1, pyref.L, 2**255 - 1, 2**252, 0x8888888888888888888888888888888888888888888888888888888888888888 >> 1,
            0x7777777777777777777777777777777777777777777777777777777777777777]
    for k, e in enumerate(edge):
        a[k], b[k] = e, edge[-1 - k]
        a[100 + k], b[100 + k] = e, 0
        a[200 + k], b[200 + k] = 0, e
    a[4199], b[4199] = pyref.L - 1, pyref.L - 1              # last lane of a partial block
    ab, bb = b"".join(b32(x) for x in a), b"".join(b32(x) for x in b)
    rc_o, want = oracle.ristretto_double_base_batch(ab, bb, G, Hc)
    eng.set_option("double_base_comb", variant)
    try:
        rc, got = eng.ristretto_double_base_batch(ab, bb, G, Hc, n)
        assert rc == rc_o == 0
        bad = [i for i in range(n) if got[32 * i:32 * i + 32] != want[32 * i:32 * i + 32]]
        assert not bad, bad[:10]
        rc, _ = eng.ristretto_double_base_batch(ab, bb, b32(1), Hc, n)   # undecodable base -> None
        assert rc == 1
    finally:
        eng.set_option("double_base_comb", 1)


def test_ristretto_double_base_streamed_pieces(eng, oracle, kat):
    """>= 2^16 pairs are streamed in pieces over two streams (comb: two full waves = 2 * SMs * 384 pairs per piece;
    Straus: 8 equal pieces): both kernels must agree everywhere and match the oracle around every piece boundary
    and on a random sample."""
    import numpy as np
    G = bytes.fromhex(kat["constants"]["RISTRETTO_BASEPOINT_COMPRESSED"]["hex"])
    Gp = oracle.ristretto_decompress(G)
    Hc = oracle.ristretto_compress(oracle.scalarmul(b32(pyref.labelled_scalar(b"dalek-b200/H", 1, 0)), Gp))
    n = 250001
    rng = np.random.Generator(np.random.PCG64(58))
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 31] &= 0x7F
    b = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); b[:, 31] &= 0x0F
    outs = []
    for variant in (0, 1):
        eng.set_option("double_base_comb", variant)
        out = np.zeros(32 * n, dtype=np.uint8)
        rc, _ = eng.ristretto_double_base_batch(a, b, G, Hc, n, out=out)
        assert rc == 0
        outs.append(out.reshape(n, 32))
    eng.set_option("double_base_comb", 1)
    assert np.array_equal(outs[0], outs[1])
    import torch
    wave2 = 2 * torch.cuda.get_device_properties(0).multi_processor_count * 384
    idx = sorted({min(n - 1, max(0, b + d)) for b in [n * k // 8 for k in range(9)] + [wave2, 2 * wave2] for d in (-1, 0, 1)}
                 | set(rng.integers(0, n, size=150).tolist()))
    rc, want = oracle.ristretto_double_base_batch(a[idx].tobytes(), b[idx].tobytes(), G, Hc)
    assert rc == 0 and outs[1][idx].tobytes() == want
    # a scalar with bit 255 set violates the Scalar invariant: rejected before any work
    a[n // 2, 31] |= 0x80
    import curve25519_dalek_b200 as pkg
    with pytest.raises(pkg.EngineError):
        eng.ristretto_double_base_batch(a, b, G, Hc, n, out=np.zeros(32 * n, dtype=np.uint8))


def test_ristretto_vartime_msm(eng, oracle, kat):
    rnd = random.Random(56)
    G = oracle.ristretto_decompress(bytes.fromhex(kat["constants"]["RISTRETTO_BASEPOINT_COMPRESSED"]["hex"]))
    n = 260
    xs = [rnd.randrange(pyref.L) for _ in range(n)]
    ts = [rnd.randrange(pyref.L) for _ in range(n)]
    pts = [oracle.ristretto_compress(oracle.scalarmul(b32(t), G)) for t in ts]
    want = oracle.ristretto_compress(oracle.scalarmul(b32(sum(x * t for x, t in zip(xs, ts)) % pyref.L), G))
    rc, got = eng.ristretto_vartime_msm(b"".join(b32(x) for x in xs), b"".join(pts), n)
    assert rc == 0 and got == want
    pts[3] = b32(1)                                       # negative s: invalid encoding -> None
    rc, _ = eng.ristretto_vartime_msm(b"".join(b32(x) for x in xs), b"".join(pts), n)
    assert rc == 1


def test_mul_base_batch(eng, oracle, kat):
    rnd = random.Random(57)
    n = 200
    xs = [rnd.randrange(2**255) for _ in range(n)]
    xs[0], xs[1], xs[2], xs[3] = 0, 1, pyref.L - 1, 2**256 - 1
    xs[4] = int.from_bytes(bytes.fromhex(kat["edwards"]["A_SCALAR"]["hex"]), "little")
    limbs, comp = eng.mul_base_batch(b"".join(b32(x) for x in xs), n)
    B = oracle.basepoint()
    for i in range(n):
        want = oracle.compress(oracle.scalarmul(b32(xs[i] % pyref.L), B))
        assert comp[32 * i:32 * i + 32] == want
        assert oracle.compress(oracle.p3_from_limbs(list(limbs[20 * i:20 * i + 20]))) == want
    assert comp[32 * 4:32 * 5] == bytes.fromhex(kat["edwards"]["A_TIMES_BASEPOINT"]["hex"])


def test_gpu_signer_matches_rfc8032(eng, oracle):
    import json, os
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "ed25519_testvectors.json")) as f:
        tv = json.load(f)["vectors"]
    H = bytes.fromhex
    seeds = b"".join(H(v["seed"]) for v in tv)
    msgs = [H(v["msg"]) for v in tv]
    offs = np.zeros(len(tv) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(m) for m in msgs])
    flat = np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy()
    pks, sigs = eng.sign_batch_flat(seeds, flat, offs, len(tv))
    for i, v in enumerate(tv):                            # TESTVECTORS: byte-exact keys and signatures
        assert pks[32 * i:32 * i + 32] == H(v["pk"])
        assert sigs[64 * i:64 * i + 64] == H(v["sig"])
    from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PrivateKey
    rnd = random.Random(58)
    seed, msg = rnd.randbytes(32), rnd.randbytes(59)
    pk, sg = eng.sign_batch_flat(seed, np.frombuffer(msg, dtype=np.uint8).copy(), np.array([0, 59], dtype=np.uint64), 1)
    assert sg == Ed25519PrivateKey.from_private_bytes(seed).sign(msg)
