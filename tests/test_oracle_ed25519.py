"""Pin the oracle's Ed25519 layer: the reference's TESTVECTORS and VALIDATIONVECTORS fixtures,
RFC 8032 signing against the `cryptography` package, the Merlin/STROBE framing against the
public merlin conformance vector, and verify_batch verdict behaviour (E/batch.rs).  CPU only."""
import ctypes as C
import hashlib
import json
import os
import random

import pytest

import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = bytes.fromhex
OK, VERIFY, ARRAYLEN, SCALARFMT, POINTDEC = 0, 1, 2, 3, 4


@pytest.fixture(scope="module")
def testvectors():
    with open(os.path.join(ROOT, "tests", "golden", "ed25519_testvectors.json")) as f:
        return json.load(f)["vectors"]


@pytest.fixture(scope="module")
def validation():
    with open(os.path.join(ROOT, "tests", "golden", "ed25519_validation.json")) as f:
        return json.load(f)["vectors"]


def test_testvectors_sign_and_verify(oracle, testvectors):
    """ed25519-dalek/tests/ed25519.rs:45-98 against TESTVECTORS (sign.input)."""
    for v in testvectors:
        seed, pk, msg, sig = H(v["seed"]), H(v["pk"]), H(v["msg"]), H(v["sig"])
        assert oracle.public_key(seed) == pk
        assert oracle.sign(msg, seed) == sig
        assert oracle.verify(msg, sig, pk) == OK
        assert oracle.verify(msg, sig, pk, strict=True) == OK
        bad = bytearray(sig); bad[5] ^= 1
        assert oracle.verify(msg, bytes(bad), pk) != OK


def test_sign_vs_cryptography(oracle):
    from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PrivateKey
    from cryptography.hazmat.primitives import serialization
    rnd = random.Random(8)
    for _ in range(16):
        seed, msg = rnd.randbytes(32), rnd.randbytes(rnd.randrange(0, 200))
        sk = Ed25519PrivateKey.from_private_bytes(seed)
        pk = sk.public_key().public_bytes(serialization.Encoding.Raw, serialization.PublicFormat.Raw)
        assert oracle.public_key(seed) == pk
        assert oracle.sign(msg, seed) == sk.sign(msg)


VERIFY_ALLOWED = {"low_order_A", "low_order_R", "non_canonical_A", "low_order_component_A",
                  "low_order_component_R", "reencoded_k"}
STRICT_ALLOWED = {"low_order_component_A", "low_order_component_R"}


def test_validation_vectors(oracle, validation):
    """ed25519-dalek/tests/validation_criteria.rs:8-23, :134-170 (914 C2SP/CCTV vectors)."""
    for v in validation:
        pk, sig, msg, flags = H(v["key"]), H(v["sig"]), v["msg"].encode(), set(v["flags"])
        ok = oracle.verify(msg, sig, pk) == OK
        assert ok == flags.issubset(VERIFY_ALLOWED), v["number"]
        ok = oracle.verify(msg, sig, pk, strict=True) == OK
        assert ok == flags.issubset(STRICT_ALLOWED), v["number"]


def test_strobe_merlin_conformance(oracle):
    """Public merlin crate vectors (strobe.rs `test_conformance`, transcript.rs
    `test_simple` equivalents).  These pin the STROBE-128 framing; the reference tree holds
    no golden transcript output of its own."""
    lib = oracle.lib
    import oracle_lib

    class Strobe(C.Structure):
        _fields_ = [("st", C.c_uint8 * 200), ("pos", C.c_uint8), ("pos_begin", C.c_uint8), ("cur_flags", C.c_uint8)]
    s = Strobe()
    proto = b"Conformance Test Protocol"
    lib.strobe128_new(C.byref(s), oracle_lib.buf(proto), C.c_size_t(len(proto)))
    msg = bytes([99]) * 1024
    lib.strobe128_meta_ad(C.byref(s), oracle_lib.buf(b"ms"), C.c_size_t(2), 0)
    lib.strobe128_meta_ad(C.byref(s), oracle_lib.buf(b"g"), C.c_size_t(1), 1)
    lib.strobe128_ad(C.byref(s), oracle_lib.buf(msg), C.c_size_t(1024), 0)
    prf1 = (C.c_uint8 * 32)()
    lib.strobe128_meta_ad(C.byref(s), oracle_lib.buf(b"prf"), C.c_size_t(3), 0)
    lib.strobe128_prf(C.byref(s), prf1, C.c_size_t(32), 0)
    assert bytes(prf1).hex() == "b48e645ca17c667fd5206ba57a6a228d72d8e1903814d3f17f622996d7cfefb0"
    lib.strobe128_meta_ad(C.byref(s), oracle_lib.buf(b"key"), C.c_size_t(3), 0)
    lib.strobe128_key(C.byref(s), prf1, C.c_size_t(32), 0)
    prf2 = (C.c_uint8 * 32)()
    lib.strobe128_meta_ad(C.byref(s), oracle_lib.buf(b"prf"), C.c_size_t(3), 0)
    lib.strobe128_prf(C.byref(s), prf2, C.c_size_t(32), 0)
    assert bytes(prf2).hex() == "07e45cce8078cee259e3e375bb85d75610e2d1e1201c5f645045a194edd49ff8"

    # merlin README / transcript.rs simple vector
    class T(C.Structure):
        _fields_ = [("s", Strobe)]
    t = T()
    lib.merlin_new(C.byref(t), oracle_lib.buf(b"test protocol"), C.c_size_t(13))
    lib.merlin_append_message(C.byref(t), oracle_lib.buf(b"some label"), C.c_size_t(10),
                              oracle_lib.buf(b"some data"), C.c_size_t(9))
    ch = (C.c_uint8 * 32)()
    lib.merlin_challenge_bytes(C.byref(t), oracle_lib.buf(b"challenge"), C.c_size_t(9), ch, C.c_size_t(32))
    assert bytes(ch).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def make_batch(oracle, n, seed=0, msg_len=59):
    rnd = random.Random(seed)
    msgs, sigs, pks = [], [], []
    for i in range(n):
        sk = rnd.randbytes(32)
        m = rnd.randbytes(msg_len if msg_len is not None else rnd.randrange(0, 300))
        msgs.append(m); pks.append(oracle.public_key(sk)); sigs.append(oracle.sign(m, sk))
    return msgs, sigs, pks


@pytest.mark.parametrize("n", [0, 1, 7, 64, 95, 96, 200])
def test_verify_batch_valid(oracle, n):
    """ed25519-dalek/tests/ed25519.rs:458-484 (verify_batch_seven_signatures) and more sizes:
    n=64 -> Straus (129 terms), n>=95 -> Pippenger."""
    msgs, sigs, pks = make_batch(oracle, n, seed=n, msg_len=None)
    assert oracle.verify_batch(msgs, sigs, pks) == OK
    if n:
        assert oracle.verify_batch(msgs, sigs, pks, chunk=3) == OK


def test_verify_batch_negative_controls(oracle):
    msgs, sigs, pks = make_batch(oracle, 9, seed=99)
    # flipped bit in s (still canonical) -> Verify
    bad = list(sigs); b = bytearray(bad[4]); b[33] ^= 1; bad[4] = bytes(b)
    assert oracle.verify_batch(msgs, bad, pks) == VERIFY
    # flipped bit in a message -> Verify
    bm = list(msgs); bm[2] = bytes([bm[2][0] ^ 1]) + bm[2][1:]
    assert oracle.verify_batch(bm, sigs, pks) == VERIFY
    # R not on curve -> Verify (E/batch.rs:235,244)
    bad = list(sigs); bad[0] = (2).to_bytes(32, "little") + bad[0][32:]
    assert oracle.verify_batch(msgs, bad, pks) == VERIFY
    # non-canonical s (s + l) -> ScalarFormat (E/batch.rs:208-211)
    s = int.from_bytes(sigs[1][32:], "little") + pyref.L
    bad = list(sigs); bad[1] = sigs[1][:32] + s.to_bytes(32, "little")
    assert oracle.verify_batch(msgs, bad, pks) == SCALARFMT
    # bad public key -> PointDecompression (VerifyingKey::from_bytes)
    bk = list(pks); bk[3] = (2).to_bytes(32, "little")
    assert oracle.verify_batch(msgs, sigs, bk) == POINTDEC
    # swapped signatures -> Verify
    sw = list(sigs); sw[0], sw[1] = sw[1], sw[0]
    assert oracle.verify_batch(msgs, sw, pks) == VERIFY


def test_verify_batch_zs_depend_on_inputs_and_chunking(oracle):
    msgs, sigs, pks = make_batch(oracle, 8, seed=5)
    rc, z1 = oracle.verify_batch(msgs, sigs, pks, want_zs=True)
    rc, z2 = oracle.verify_batch(msgs, sigs, pks, want_zs=True)
    assert rc == OK and z1 == z2 and len(set(z1[16 * i:16 * i + 16] for i in range(8))) == 8
    # chunked transcripts: chunk k's z values equal those of a stand-alone batch over that chunk
    rc, zc = oracle.verify_batch(msgs, sigs, pks, chunk=4, want_zs=True)
    rc, za = oracle.verify_batch(msgs[:4], sigs[:4], pks[:4], want_zs=True)
    rc, zb = oracle.verify_batch(msgs[4:], sigs[4:], pks[4:], want_zs=True)
    assert zc == za + zb and zc != z1


def test_mixed_order_batches_have_z_dependent_verdicts(oracle):
    """tests/torsion_cases.py (inputs of the GPU parity test for small-order components): over its seeds the reference
    algorithm returns both Ok and Verify, and cutting the transcript into 64-signature chunks changes some verdicts --
    which is why the engine's default is the reference's single transcript."""
    import torsion_cases
    verdicts, disagreements = set(), 0
    for trial in range(10):
        msgs, sigs, pks = torsion_cases.make_batch(oracle, 100, seed=1000 * 100 + trial)
        whole = oracle.verify_batch(msgs, sigs, pks)
        verdicts.add(whole)
        disagreements += whole != oracle.verify_batch(msgs, sigs, pks, chunk=64)
        # every signature is individually INVALID under `verify` only when its own defect is non-zero: the clean ones pass
        assert oracle.verify(msgs[1], sigs[1], pks[1]) in (0, 1)
    assert verdicts == {0, 1} and disagreements > 0


def test_small_order_part_of_a_batch_equation_from_scalars_mod_8(oracle):
    """The identity behind the engine's per-batch small-order test (csrc/batch.cu, k_batch_torsion): with E the value of the
    batch equation (batch.rs:240-244), S = sum (z_i mod 8) R_i + sum ((z_i h_i mod l) mod 8) A_i and the digits
    k = a + 3 b (a, b in {-1, 0, 1}) the kernel uses,
        [l] S == identity   <=>   the small-order part of E is zero,
    and E == identity (the reference's verdict) <=> [8] E == identity (prime-order part) and [l] S == identity.
    Checked on the torsion batches of tests/torsion_cases.py, where both outcomes occur, with the oracle's own z_i."""
    import torsion_cases
    L = pyref.L
    lb = L.to_bytes(32, "little")
    AB = {0: (0, 0), 1: (1, 0), 2: (-1, 1), 3: (0, 1), 4: (1, 1), 5: (0, -1), 6: (1, -1), 7: (-1, 0)}   # k = a + 3 b mod 8
    assert all((a + 3 * b) % 8 == k for k, (a, b) in AB.items())
    seen = set()
    for trial in range(12):
        n = 24
        msgs, sigs, pks = torsion_cases.make_batch(oracle, n, seed=77000 + trial)
        rc, zs = oracle.verify_batch(msgs, sigs, pks, want_zs=True)
        assert rc in (0, 1)
        B = oracle.basepoint()
        E, B1, B3 = oracle.identity(), oracle.identity(), oracle.identity()
        sum_zs = 0
        for i in range(n):
            z = int.from_bytes(zs[16 * i:16 * i + 16], "little")
            R, A = oracle.decompress(sigs[i][:32]), oracle.decompress(pks[i])
            assert R is not None and A is not None
            h = int.from_bytes(hashlib.sha512(sigs[i][:32] + pks[i] + msgs[i]).digest(), "little") % L
            zh = z * h % L
            sum_zs = (sum_zs + z * int.from_bytes(sigs[i][32:], "little")) % L
            E = oracle.add(E, oracle.scalarmul(z.to_bytes(32, "little"), R))
            E = oracle.add(E, oracle.scalarmul(zh.to_bytes(32, "little"), A))
            for k, P in ((z % 8, R), (zh % 8, A)):
                a, b = AB[k]
                if a:
                    B1 = oracle.add(B1, P) if a > 0 else oracle.sub(B1, P)
                if b:
                    B3 = oracle.add(B3, P) if b > 0 else oracle.sub(B3, P)
        E = oracle.add(E, oracle.scalarmul(((L - sum_zs) % L).to_bytes(32, "little"), B))
        assert oracle.is_identity(E) == (rc == 0)                         # the reference's verdict is this equation
        S = oracle.add(B1, oracle.add(oracle.double(B3), B3))             # B1 + 3 B3
        lS = oracle.scalarmul(lb, S)
        small_zero = oracle.is_identity(lS)
        prime_zero = oracle.is_identity(oracle.mul_by_pow_2(E, 3))
        lE = oracle.scalarmul(lb, E)
        assert small_zero == oracle.is_identity(lE)                       # [l] S and [l] E are the same small-order test
        assert (rc == 0) == (prime_zero and small_zero)
        seen.add((prime_zero, small_zero))
    assert (True, True) in seen and (True, False) in seen
