"""Host build of the device field/point headers (fe.cuh, ge.cuh) with limb-bound assertions,
compared operation by operation with the CPU oracle.  CPU only; the host build is test
infrastructure, not a fallback of the product."""
import ctypes as C
import os
import random
import subprocess

import pytest

import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    src = os.path.join(ROOT, "tests", "host", "fe_host_check.cpp")
    so = os.path.join(ROOT, "tests", "host", "libfehost.so")
    deps = [src] + [os.path.join(ROOT, "curve25519_dalek_b200", "csrc", f) for f in ("fe.cuh", "ge.cuh", "constants.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


def b32(x):
    return (x % 2**256).to_bytes(32, "little")


def call(fn, *args, nout=1):
    outs = [(C.c_uint8 * 32)() for _ in range(nout)]
    fn(*outs, *[(C.c_uint8 * 32).from_buffer_copy(a) for a in args])
    r = [int.from_bytes(bytes(o), "little") for o in outs]
    return r[0] if nout == 1 else r


EDGE = [0, 1, 2, 19, pyref.p - 1, pyref.p, pyref.p + 1, 2**255 - 1, 2**255 - 19 - 1, (1 << 255) - 20,
        2**26 - 1, 2**51, (2**255 - 1) ^ (2**128 - 1)]


def test_field_ops(host):
    rnd = random.Random(1)
    vals = EDGE + [rnd.randrange(2**255) for _ in range(300)]
    p = pyref.p
    for i, x in enumerate(vals):
        y = vals[(i * 7 + 3) % len(vals)]
        assert call(host.h_fe_mul, b32(x), b32(y)) == x * y % p
        assert call(host.h_fe_sq, b32(x)) == x * x % p
        assert call(host.h_fe_add, b32(x), b32(y)) == (x + y) % p
        assert call(host.h_fe_sub, b32(x), b32(y)) == (x - y) % p
        assert call(host.h_fe_chain, b32(x), b32(y)) == ((x * y - x) ** 2 + y) * (x - y) % p
    for x in vals[:40]:
        assert call(host.h_fe_invert, b32(x)) == pow(x % p, p - 2, p)
        assert call(host.h_fe_pow_p58, b32(x)) == pow(x % p, (p - 5) // 8, p)


def test_decompress_matches_oracle(host, oracle):
    rnd = random.Random(2)
    cases = [b32(v) for v in EDGE] + [rnd.randbytes(32) for _ in range(200)]
    B = oracle.basepoint()
    cases += [oracle.compress(oracle.scalarmul(b32(rnd.randrange(pyref.L)), B)) for _ in range(20)]
    nvalid = 0
    for s in cases:
        ox, oy = (C.c_uint8 * 32)(), (C.c_uint8 * 32)()
        ok = host.h_decompress(ox, oy, (C.c_uint8 * 32).from_buffer_copy(s))
        P = oracle.decompress(s)
        assert bool(ok) == (P is not None)
        if P is not None:
            nvalid += 1
            assert bytes(ox) == oracle.fe_to_bytes(P.X) and bytes(oy) == oracle.fe_to_bytes(P.Y)
    assert nvalid > 50


def test_point_ops_match_oracle(host, oracle):
    rnd = random.Random(3)
    B = oracle.basepoint()
    ident = oracle.compress(oracle.identity())
    torsion = [b32(0), b32(pyref.p - 1)]          # (x,0)-type 4-torsion and (0,-1)
    pts = [oracle.compress(oracle.scalarmul(b32(rnd.randrange(pyref.L)), B)) for _ in range(12)]
    pts += [ident] + torsion
    for i, Pc in enumerate(pts):
        Qc = pts[(i * 5 + 1) % len(pts)]
        P, Q = oracle.decompress(Pc), oracle.decompress(Qc)
        if P is None or Q is None:
            continue
        k = 1 + i % 7
        outs = [(C.c_uint8 * 32)() for _ in range(6)]
        rc = host.h_point_ops(*outs, (C.c_uint8 * 32).from_buffer_copy(Pc), (C.c_uint8 * 32).from_buffer_copy(Qc), k)
        assert rc == 1
        o_add, o_sub, o_madd, o_msub, o_dbl, o_pow = [bytes(o) for o in outs]
        assert o_add == oracle.compress(oracle.add(P, Q))
        assert o_sub == oracle.compress(oracle.sub(P, Q))
        pp = oracle.add(oracle.double(P), Q)
        assert o_madd == oracle.compress(oracle.add(pp, Q))
        assert o_msub == oracle.compress(oracle.double(P))
        assert o_dbl == oracle.compress(oracle.double(pp))
        assert o_pow == oracle.compress(oracle.mul_by_pow_2(pp, k))
        assert host.h_is_identity_of_diff((C.c_uint8 * 32).from_buffer_copy(Pc)) == 1


def test_limbs51_conversion(host, oracle):
    rnd = random.Random(4)
    B = oracle.basepoint()
    for _ in range(20):
        P = oracle.scalarmul(b32(rnd.randrange(pyref.L)), B)
        P = oracle.add(P, oracle.double(P))
        limbs = oracle.p3_limbs(P)
        # unreduced limbs up to 2^54 are legal inputs (u64/field.rs:27-43)
        fat = list(limbs)
        fat[0] += 7 * (2**51 - 19); fat[1] += 7 * (2**51 - 1); fat[2] += 7 * (2**51 - 1)
        fat[3] += 7 * (2**51 - 1); fat[4] += 7 * (2**51 - 1)
        out = (C.c_uint64 * 20)()
        host.h_limbs51_roundtrip(out, (C.c_uint64 * 20)(*fat))
        for c in range(4):
            f = oracle_fe(oracle, limbs[5 * c:5 * c + 5])
            want = int.from_bytes(oracle.fe_to_bytes(f), "little")
            got = sum(int(out[5 * c + i]) << (51 * i) for i in range(5))
            assert got == want and all(int(out[5 * c + i]) < 2**51 for i in range(5))


def oracle_fe(oracle, limbs):
    import oracle_lib
    f = oracle_lib.Fe()
    for i, v in enumerate(limbs):
        f.v[i] = v
    return f


def test_ristretto_encode_decode(host, oracle, kat):
    rnd = random.Random(9)
    encs = [bytes.fromhex(h) for h in kat["ristretto"]["SMALL_MULTIPLES"]["hex"]]
    B = oracle.ristretto_decompress(encs[1])
    for _ in range(20):
        encs.append(oracle.ristretto_compress(oracle.scalarmul(b32(rnd.randrange(pyref.L)), B)))
    bad = [b32(1), b32(pyref.p), b"\xff" * 32, b32(2**255 + 2), rnd.randbytes(32), rnd.randbytes(32)]
    for e in encs + bad:
        o1, o2 = (C.c_uint8 * 32)(), (C.c_uint8 * 32)()
        ok = host.h_ristretto_roundtrip(o1, o2, (C.c_uint8 * 32).from_buffer_copy(e))
        P = oracle.ristretto_decompress(e)
        assert bool(ok) == (P is not None)
        if P is not None:
            assert bytes(o1) == e
            assert bytes(o2) == oracle.ristretto_compress(oracle.add(oracle.double(P), P))


def test_fe64_model(host, oracle):
    """Host model of the FP64-pipe field (fe64.cuh): exact-product arithmetic, balanced carries and the
    operand rule (asserted inside), against big integers and the oracle."""
    rnd = random.Random(10)
    p = pyref.p
    vals = EDGE + [rnd.randrange(2**255) for _ in range(200)]
    for i, x in enumerate(vals):
        y = vals[(i * 5 + 1) % len(vals)]
        assert call(host.h_fe64_mul, b32(x), b32(y)) == (x * y * y + x - y) * x % p
        assert call(host.h_fe64_sq, b32(x), b32(y)) == pow(pow(x, 4, p) * y, 2, p)
        if i < 60:
            assert call(host.h_fe64_pow_p58, b32(x)) == pow(x, (p - 5) // 8, p)
            assert call(host.h_fe64_invert, b32(x)) == pow(x, p - 2, p)
    B = oracle.basepoint()
    ident = oracle.compress(oracle.identity())
    for trial in range(6):
        n = 40
        pts = [oracle.compress(oracle.scalarmul(b32(rnd.randrange(pyref.L)), B)) for _ in range(n)]
        if trial == 0:
            pts[3] = ident; pts[4] = b32(0); pts[5] = b32(p - 1)
        negs = bytes(rnd.randrange(2) for _ in range(n))
        out = (C.c_uint8 * 32)()
        for k in (0, 1, 2, 7) if trial < 2 else ():                     # doubling chain on the FP64 field
            for pt in pts[:6]:
                assert host.h_ge64_dbl_chain(out, pt, k) == 1
                want = oracle.scalarmul(b32(3 * (2**k + 1)), oracle.decompress(pt))
                assert bytes(out) == oracle.compress(want)
        assert host.h_ge64_chain(out, b"".join(pts), negs, n) == 1
        acc = oracle.identity()
        for k in range(n):
            q = oracle.decompress(pts[k])
            if k % 2 == 0:
                q = oracle.add(oracle.double(q), q)
            acc = oracle.sub(acc, q) if negs[k] else oracle.add(acc, q)
        assert bytes(out) == oracle.compress(acc)
