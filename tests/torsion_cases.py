"""Batches of Ed25519 signatures whose R or public key carries a small-order component (test input synthesis).

For such inputs the batch equation of ed25519_dalek::verify_batch (batch.rs:240-250, no cofactor multiplication) is
left with a pure torsion defect, so Ok / Verify depends on the coefficients z_i modulo 8 (components in R) and on
(z_i h_i mod l) modulo 8 (components in A): only the reference's own transcript gives the reference's verdict."""
import hashlib
import random

import pyref

T8 = bytes.fromhex("c7176a703d4dd84fba3c0b760d10670f2a2053fa2c39ccc64ec7fd7792ac037a")    # order 8 (u64/constants.rs:196-340)


def torsion_points(orc):
    """[T, 2T, ..., 7T] as oracle points (orders 8, 4, 8, 2, 8, 4, 8)."""
    t8 = orc.decompress(T8)
    assert t8 is not None
    pts = [t8]
    for _ in range(6):
        pts.append(orc.add(pts[-1], t8))
    assert orc.is_identity(orc.add(pts[-1], t8))
    return pts


def sign_with_torsion(orc, rnd, msg, a, A_enc, t_R):
    """s = r + H(R' || A || M) a with R' = rB + t_R (RFC 8032 signing with a shifted commitment)."""
    r = rnd.randrange(pyref.L)
    R = orc.scalarmul(r.to_bytes(32, "little"), orc.basepoint())
    if t_R is not None:
        R = orc.add(R, t_R)
    R_enc = orc.compress(R)
    h = int.from_bytes(hashlib.sha512(R_enc + A_enc + msg).digest(), "little") % pyref.L
    return R_enc + ((r + h * a) % pyref.L).to_bytes(32, "little")


def make_batch(orc, n, seed):
    """(msgs, sigs, pks): three keys -- one clean, one with an order-2 component, one with a random component -- sign
    n messages in turn; every fifth R carries the order-2 point, one R a random small-order point (odd seeds)."""
    tors = torsion_points(orc)
    rnd = random.Random(seed)
    secrets = [rnd.randrange(pyref.L) for _ in range(3)]
    mild = seed % 2 == 0                                     # even seeds: order-2 components only (Ok with probability ~1/2)
    key_tors = [None, tors[3], tors[3] if mild else tors[rnd.randrange(7)]]
    A_enc = []
    for a, t in zip(secrets, key_tors):
        p = orc.scalarmul(a.to_bytes(32, "little"), orc.basepoint())
        A_enc.append(orc.compress(p if t is None else orc.add(p, t)))
    msgs, sigs, pks = [], [], []
    for i in range(n):
        k = i % 3
        m = rnd.randbytes(rnd.randrange(0, 80))
        t_R = tors[3] if i % 5 == 0 else (tors[rnd.randrange(7)] if (i == 7 and not mild) else None)
        msgs.append(m); pks.append(A_enc[k]); sigs.append(sign_with_torsion(orc, rnd, m, secrets[k], A_enc[k], t_R))
    return msgs, sigs, pks
