"""The product's 4-lane FP64 point code (csrc/warp4_f64.cuh: the Horner pass of k_combine; csrc/straus_vt.cuh: NAF,
table of odd multiples and main loop of the vartime Straus path) executed on the CPU by an emulated warp of 32 host
threads (tests/host/w4_host_check.cpp) with the operand-rule assertions of the host field model switched on, against
the oracle.  CPU only: the emulation is test infrastructure, not a fallback of the product."""
import ctypes as C
import os
import random
import subprocess

import pytest

import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def w4():
    src = os.path.join(ROOT, "tests", "host", "w4_host_check.cpp")
    so = os.path.join(ROOT, "tests", "host", "libw4host.so")
    csrc = os.path.join(ROOT, "curve25519_dalek_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("fe.cuh", "fe64.cuh", "ge.cuh", "ge64.cuh", "warp4_f64.cuh", "straus_vt.cuh", "transcript_warp.cuh", "constants.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, src, "-lpthread"])
    return C.CDLL(so)


def _points(oracle, rnd, n, special=True):
    B = oracle.basepoint()
    pts = [oracle.scalarmul(rnd.randrange(pyref.L).to_bytes(32, "little"), B) for _ in range(n)]
    if special and n >= 4:
        pts[1] = oracle.identity()
        pts[2] = oracle.decompress((pyref.p - 1).to_bytes(32, "little"))       # (0, -1), order 2
        pts[3] = oracle.decompress((0).to_bytes(32, "little"))                 # order 4
    return pts


@pytest.mark.parametrize("ranks,nwin,c", [(1, 17, 16), (3, 4, 20), (8, 5, 4), (2, 1, 7)])
def test_horner_pass_matches_oracle(w4, oracle, ranks, nwin, c):
    """w4f_horner (k_combine): sum_w 2^(c w) * sum_r W[r][w], doublings and additions on four lanes."""
    rnd = random.Random(ranks * 1000 + nwin)
    pts = _points(oracle, rnd, ranks * nwin)
    want = oracle.identity()
    for w in range(nwin - 1, -1, -1):
        want = oracle.mul_by_pow_2(want, c) if w != nwin - 1 else want
        for r in range(ranks):
            want = oracle.add(want, pts[r * nwin + w])
    out = (C.c_uint8 * 32)()
    buf = b"".join(oracle.compress(p) for p in pts)
    assert w4.h_w4f_horner(out, buf, ranks, nwin, c) == 1
    assert bytes(out) == oracle.compress(want)


def test_device_naf_is_the_reference_naf(w4, oracle):
    rnd = random.Random(5)
    for t in range(300):
        v = [0, 1, 2**255 - 1, 2**256 - 1, pyref.L - 1][t] if t < 5 else (rnd.randrange(2**255) if t % 3 else rnd.randrange(2**256))
        sb = v.to_bytes(32, "little")
        out = (C.c_int8 * 264)()
        w4.h_naf5(out, sb)
        d = list(out)
        assert sum(x << i for i, x in enumerate(d)) == v
        assert all(x == 0 or (x % 2 and -16 < x < 16) for x in d)
        if v < 2**255:                                            # reference-legal scalars: exactly scalar.rs:955-1007
            assert d[:256] == oracle.naf(sb, 5) and not any(d[256:])


@pytest.mark.parametrize("n", [1, 8, 9])
def test_straus_vartime_matches_oracle(w4, oracle, n):
    """straus_vt.cuh end to end on the emulated warp(s): NafLookupTable5 tables, main loop with per-group accumulators,
    sum of the groups -- against the oracle's vartime Straus (straus.rs:159-200)."""
    rnd = random.Random(40 + n)
    pts = _points(oracle, rnd, n, special=n >= 7)
    scalars = [rnd.randrange(pyref.L).to_bytes(32, "little") for _ in range(n)]
    if n >= 7:
        scalars[0] = (0).to_bytes(32, "little")
        scalars[4] = (1).to_bytes(32, "little")
        scalars[5] = (pyref.L - 1).to_bytes(32, "little")
        scalars[6] = (2**255 - 1).to_bytes(32, "little")
    want = oracle.compress(oracle.msm("straus_vartime", scalars, pts))
    out = (C.c_uint8 * 32)()
    assert w4.h_straus_vartime(out, b"".join(scalars), b"".join(oracle.compress(p) for p in pts), n) == 1
    assert bytes(out) == want


def test_merlin_prefix_constant(w4, oracle):
    """The state after Transcript::new(b"ed25519 batch verification") (transcript.rs:54-61) that transcript_warp.cuh
    holds as a constant, recomputed with the oracle's STROBE / Merlin code."""
    class Strobe(C.Structure):
        _fields_ = [("st", C.c_uint8 * 200), ("pos", C.c_uint8), ("pos_begin", C.c_uint8), ("cur_flags", C.c_uint8)]
    t = Strobe()
    lab = b"ed25519 batch verification"
    oracle.lib.merlin_new(C.byref(t), lab, C.c_size_t(len(lab)))
    lanes = (C.c_uint64 * 25)(); pos = C.c_uint32(); pb = C.c_uint32()
    w4.h_merlin_prefix(lanes, C.byref(pos), C.byref(pb))
    st = bytes(t.st)
    assert list(lanes) == [int.from_bytes(st[8 * i:8 * i + 8], "little") for i in range(25)]
    assert (pos.value, pb.value) == (t.pos, t.pos_begin)


@pytest.mark.parametrize("n", [1, 2, 9, 53])
def test_warp_transcript_draws_the_reference_coefficients(w4, oracle, n):
    """merlin_zs_warp (the 25-lane Keccak and the closed-form byte stream of the transcript) on the emulated warp: every
    z_i equals what the oracle's Merlin transcript draws (batch.rs:168-222).  n = 53 ends exactly on a rate-block
    boundary (no forced permutation before the KEY operation)."""
    rnd = random.Random(n)
    msgs = [rnd.randbytes(rnd.randrange(0, 40)) for _ in range(n)]
    sks = [rnd.randbytes(32) for _ in range(n)]
    pks = [oracle.public_key(s) for s in sks]
    sigs = [oracle.sign(m, s) for m, s in zip(msgs, sks)]
    rc, want = oracle.verify_batch(msgs, sigs, pks, want_zs=True)
    assert rc == 0
    import hashlib
    hr = b"".join(hashlib.sha512(sigs[i][:32] + pks[i] + msgs[i]).digest() for i in range(n))
    out = (C.c_uint8 * (16 * n))()
    w4.h_merlin_zs(out, hr, b"".join(sigs), C.c_uint64(n))
    assert bytes(out) == want


def test_20_lane_field_multiplication_at_the_operand_rule_extremes(w4):
    """w20_mul / w20_carry (one limb per lane: the doublings of the Horner pass) against the single-thread model: random
    balanced limbs and the corners of the operand rule -- scales 3 x 2, 2 x 2, 1 x 4 + the 2^15 slack, every limb at +- the
    bound, single non-zero limbs -- so that every column sum and every carry of the parallel round sees its extreme."""
    rnd = random.Random(20)
    B = 1 << 50
    dbl = C.c_double * 20
    w4.h_w20_mul.argtypes = [dbl, dbl]
    w4.h_w20_carry.argtypes = [dbl]

    def limbs(scale, mode):
        lim = scale * B + (scale << 14)
        if mode == "rand":
            return [float(rnd.randrange(-lim, lim + 1)) for _ in range(5)]
        if mode == "max":
            return [float(lim)] * 5
        if mode == "min":
            return [float(-lim)] * 5
        if mode == "alt":
            return [float(lim if k % 2 == 0 else -lim) for k in range(5)]
        v = [0.0] * 5; v[rnd.randrange(5)] = float(rnd.choice((lim, -lim))); return v      # "one"

    modes = ("rand", "max", "min", "alt", "one")
    for sa, sb in ((1, 1), (2, 2), (3, 2), (2, 3), (1, 4), (1, 7), (7, 1)):
        for trial in range(40):
            a, b = [], []
            for g in range(4):
                ma, mb = (rnd.choice(modes), rnd.choice(modes)) if trial >= 8 else (modes[(trial + g) % 5], modes[(trial * 3 + g) % 5])
                a += limbs(sa, ma); b += limbs(sb, mb)
            assert w4.h_w20_mul(dbl(*a), dbl(*b)) == 1, (sa, sb, trial)
    for trial in range(200):
        a = []
        for g in range(4):
            a += limbs(rnd.choice((1, 2, 4, 64, 4096)), rnd.choice(modes))
        assert w4.h_w20_carry(dbl(*a)) == 1, trial


@pytest.mark.parametrize("k", [1, 2, 16, 65])
def test_20_lane_doublings_match_oracle(w4, oracle, k):
    rnd = random.Random(k)
    for P in _points(oracle, rnd, 6):
        out = (C.c_uint8 * 32)()
        assert w4.h_w20_dbl_n(out, oracle.compress(P), k) == 1
        assert bytes(out) == oracle.compress(oracle.mul_by_pow_2(P, k))

