"""Device-level parity of the product's field arithmetic (csrc/fe.cuh, fe64.cuh) against exact integer arithmetic
mod 2^255 - 19 -- the semantics of FieldElement51::mul / square / to_bytes
(curve25519-dalek/src/backend/serial/u64/field.rs:111-214, :454-559, :368-450).

A TEST-ONLY kernel (tests/device/fe_device_check.cu, compiled from the product headers) takes raw limb operands, so
the operands can sit AT THE EXTREMES of the limb-size rules the headers document -- limb products of 2^103 for the
FP64 exact split (fe64.cuh: __fma_rz / magic-constant rounding), scale products of 8, 32-bit limbs at scale 30 --
which canonical random inputs through the point formulas never reach.  About 1.6 million operand pairs in total."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**255 - 19
N = 1 << 18
OP_FE_MUL, OP_FE_SQ, OP_FE64_MUL, OP_FE64_SQ, OP_FE64_CARRY, OP_FE64_TO_FE, OP_FE_SUB_MUL, OP_FE64_FROM_FE = range(8)
SH25 = [0, 26, 51, 77, 102, 128, 153, 179, 204, 230]      # bit offset of limb i in radix 2^25.5
SH51 = [0, 51, 102, 153, 204]


@pytest.fixture(scope="module")
def dev():
    sys.path.insert(0, os.path.join(ROOT, "tests", "device"))
    import build as devbuild
    lib = C.CDLL(devbuild.build())
    lib.fe_device_check.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    return lib


def run(dev, op, a, b=None):
    n = a.shape[0]
    out = np.zeros((n, 8), dtype=np.uint32)
    limbs = np.zeros((n, 5), dtype=np.int64)
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b) if b is not None else None
    rc = dev.fe_device_check(op, a.ctypes.data, b.ctypes.data if b is not None else None, n, out.ctypes.data, limbs.ctypes.data)
    assert rc == 0
    return [int.from_bytes(r.tobytes(), "little") for r in out], limbs


def values(limbs, shifts):
    return [sum(int(x) << s for x, s in zip(row, shifts)) % P for row in limbs.tolist()]


def limbs25(rng, n, scale):
    """n x 10 u32 limbs: a third uniformly random below the bound, a third AT the bound, a third mixed per limb."""
    bound = np.array([int(scale * 1.01 * (1 << (25 if i & 1 else 26))) for i in range(10)], dtype=np.uint64)
    bound = np.minimum(bound, 0xffffffff)
    r = (rng.random((n, 10)) * (bound + 1)).astype(np.uint64)
    r = np.minimum(r, bound)
    third = n // 3
    r[third:2 * third] = bound
    pick = rng.random((n - 2 * third, 10)) < 0.5
    r[2 * third:] = np.where(pick, bound, r[2 * third:])
    return r.astype(np.uint32)


def limbs51(rng, n, scale, extreme_frac=2 / 3):
    """n x 5 balanced signed limbs with |limb| <= scale * 2^50 (+ scale * 2^10 slack): random, at the bound with random
    signs, and mixed."""
    bound = int(scale * (1 << 50) + scale * 1024)
    mag = (rng.random((n, 5)) * bound).astype(np.int64)
    sign = np.where(rng.random((n, 5)) < 0.5, -1, 1).astype(np.int64)
    third = n // 3
    mag[third:2 * third] = bound
    pick = rng.random((n - 2 * third, 5)) < 0.5
    mag[2 * third:] = np.where(pick, bound, mag[2 * third:])
    return mag * sign


def test_fe_mul_and_sq_at_the_limb_bounds(dev):
    """fe.cuh: fe_mul needs scale(g) <= 3.3 and scale(f) * scale(g) <= 30; fe_sq needs scale <= 2."""
    rng = np.random.default_rng(1)
    for sf, sg in ((1, 1), (9, 3.3), (30, 1), (4, 3), (2, 2)):
        f, g = limbs25(rng, N // 4, sf), limbs25(rng, N // 4, sg)
        got, _ = run(dev, OP_FE_MUL, f, g)
        vf, vg = values(f, SH25), values(g, SH25)
        assert got == [x * y % P for x, y in zip(vf, vg)], (sf, sg)
    for s in (1, 2):
        f = limbs25(rng, N // 2, s)
        got, _ = run(dev, OP_FE_SQ, f)
        assert got == [x * x % P for x in values(f, SH25)], s
    # uncarried difference and sum feeding a multiplication (the pattern of every point formula)
    f, g = limbs25(rng, N // 2, 1), limbs25(rng, N // 2, 1)
    got, _ = run(dev, OP_FE_SUB_MUL, f, g)
    vf, vg = values(f, SH25), values(g, SH25)
    assert got == [(x - y) * (x + y) % P for x, y in zip(vf, vg)]


def test_fe64_mul_exact_split_at_the_operand_rule(dev):
    """fe64.cuh: |a_i b_j| < 2^103, scale(a) * scale(b) < 8.  Products at the edge exercise the round-toward-zero split
    t = fma_rz(a, b 2^-52, 1.5 2^52) and the magic-constant low half with both signs."""
    rng = np.random.default_rng(2)
    for sa, sb in ((1, 1), (2.8, 2.8), (7.9, 1), (1, 7.9), (3.9, 2), (2, 3.9)):
        a, b = limbs51(rng, N // 4, sa), limbs51(rng, N // 4, sb)
        assert (np.abs(a).max().item() * np.abs(b).max().item()) < 2**103
        got, limbs = run(dev, OP_FE64_MUL, a, b)
        va, vb = values(a, SH51), values(b, SH51)
        assert got == [x * y % P for x, y in zip(va, vb)], (sa, sb)
        assert np.abs(limbs).max() <= (1 << 50) + 1024                # outputs have scale 1
    # tiny and sparse operands: zeros, +-1, single limbs (floor of small negative products)
    small = np.array([[0, 0, 0, 0, 0], [1, 0, 0, 0, 0], [-1, 0, 0, 0, 0], [0, 0, 0, 0, -1], [-1, -1, -1, -1, -1],
                      [(1 << 50), -(1 << 50), (1 << 50), -(1 << 50), (1 << 50)], [1, -1, 1, -1, 1]], dtype=np.int64)
    aa = np.repeat(small, len(small), axis=0)
    bb = np.tile(small, (len(small), 1))
    got, _ = run(dev, OP_FE64_MUL, aa, bb)
    assert got == [x * y % P for x, y in zip(values(aa, SH51), values(bb, SH51))]


def test_fe64_sq_carry_and_conversions(dev):
    rng = np.random.default_rng(3)
    for s in (1, 1.99):                                               # fe64_sq: 2 |a_i a_j| < 2^103, scale < 2
        a = limbs51(rng, N // 2, s)
        got, limbs = run(dev, OP_FE64_SQ, a)
        assert got == [x * x % P for x in values(a, SH51)], s
        assert np.abs(limbs).max() <= (1 << 50) + 1024
    # fe64_carry: any integer-valued input up to 2^62 (53 significant bits); output limbs |.| <= 2^50 + a few units
    n = N // 2
    mant = rng.integers(0, 1 << 53, size=(n, 5), dtype=np.int64)
    shift = rng.integers(0, 10, size=(n, 5), dtype=np.int64)          # up to 2^62
    sign = np.where(rng.random((n, 5)) < 0.5, -1, 1).astype(np.int64)
    a = (mant << shift) * sign
    a[: n // 4] = limbs51(rng, n // 4, 4)                             # the scales the point formulas produce
    got, limbs = run(dev, OP_FE64_CARRY, a)
    assert got == values(a, SH51)
    assert np.abs(limbs).max() <= (1 << 50) + (1 << 13)
    assert values(limbs, SH51) == values(a, SH51)
    # fe64_to_fe: any scale <= 4
    a = limbs51(rng, N // 2, 4)
    got, _ = run(dev, OP_FE64_TO_FE, a)
    assert got == values(a, SH51)
    # fe64_from_fe on 32-bit limbs of scale 1..4 (it canonicalises first), result scale 1
    f = limbs25(rng, N // 2, 4)
    got, limbs = run(dev, OP_FE64_FROM_FE, f)
    assert got == values(f, SH25)
    assert np.abs(limbs).max() <= (1 << 50) + 1024
