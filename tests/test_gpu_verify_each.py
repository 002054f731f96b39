"""GPU parity tests of the per-signature verifier: VerifyingKey::verify and verify_strict
(E/verifying.rs:203-219, :359-382) for many signatures at once, against the reference's own pins -- all 914
VALIDATIONVECTORS (tests/validation_criteria.rs:8-23, :134-170) and all 128 TESTVECTORS -- and the CPU oracle."""
import hashlib
import json
import os
import random

import pytest

import pyref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = bytes.fromhex
OK, VERIFY, SCALARFMT, POINTDEC = 0, 1, 3, 4
from test_oracle_ed25519 import STRICT_ALLOWED, VERIFY_ALLOWED   # the flag sets of tests/validation_criteria.rs:8-23


@pytest.fixture(scope="module")
def eng():
    import curve25519_dalek_b200 as pkg
    e = pkg.Engine(0)
    yield e
    e.close()


def flat(msgs):
    import numpy as np
    offs = np.zeros(len(msgs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(m) for m in msgs])
    return np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy(), offs


@pytest.mark.parametrize("strict", [False, True])
def test_validation_vectors(eng, oracle, strict):
    with open(os.path.join(ROOT, "tests", "golden", "ed25519_validation.json")) as f:
        vec = json.load(f)["vectors"]
    msgs = [v["msg"].encode() for v in vec]
    sigs = [H(v["sig"]) for v in vec]
    keys = [H(v["key"]) for v in vec]
    fl, offs = flat(msgs)
    rc, res = eng.verify_each_flat(fl, offs, b"".join(sigs), b"".join(keys), len(vec), strict=strict)
    allowed = STRICT_ALLOWED if strict else VERIFY_ALLOWED
    for v, r, m, s, k in zip(vec, res, msgs, sigs, keys):
        assert (r == OK) == set(v["flags"]).issubset(allowed), v["number"]
        assert r == oracle.verify(m, s, k, strict=strict), v["number"]       # the error kind too
    assert rc == (OK if all(r == OK for r in res) else VERIFY)


def test_testvectors_and_error_kinds(eng, oracle):
    with open(os.path.join(ROOT, "tests", "golden", "ed25519_testvectors.json")) as f:
        tv = json.load(f)["vectors"]
    msgs = [H(v["msg"]) for v in tv]
    sigs = [H(v["sig"]) for v in tv]
    keys = [H(v["pk"]) for v in tv]
    fl, offs = flat(msgs)
    for strict in (False, True):
        rc, res = eng.verify_each_flat(fl, offs, b"".join(sigs), b"".join(keys), len(tv), strict=strict)
        assert rc == OK and res == [OK] * len(tv)
    # one failure of every kind; everything else stays Ok
    s, k, m = list(sigs), list(keys), list(msgs)
    m[3] = m[3] + b"x"                                                    # Verify
    x = bytearray(s[10]); x[63] |= 0xf0; s[10] = bytes(x)                 # s >= l: ScalarFormat
    k[20] = (2).to_bytes(32, "little")                                    # undecodable key: PointDecompression
    x = bytearray(s[30]); x[63] |= 0xf0; s[30] = bytes(x); k[30] = (2).to_bytes(32, "little")   # both: key first
    s[40] = (2).to_bytes(32, "little") + s[40][32:]                       # undecodable R: Verify
    x = bytearray(s[50]); x[0] ^= 1; s[50] = bytes(x)                     # wrong R
    fl, offs = flat(m)
    for strict in (False, True):
        rc, res = eng.verify_each_flat(fl, offs, b"".join(s), b"".join(k), len(tv), strict=strict)
        want = [oracle.verify(m[i], s[i], k[i], strict=strict) for i in range(len(tv))]
        assert res == want and rc == VERIFY
        assert [i for i, r in enumerate(res) if r] == [3, 10, 20, 30, 40, 50]
        assert (res[3], res[10], res[20], res[30], res[40], res[50]) == (VERIFY, SCALARFMT, POINTDEC, POINTDEC, VERIFY, VERIFY)
    assert eng.verify_each_flat(fl, offs, b"", b"", 0)[0] == OK


def test_verify_each_agrees_with_batch_and_locates_failures(eng, oracle):
    """2^17 + 9 signatures (streamed in 2^16 pieces over two streams): valid ones pass; the corrupted ones are exactly
    the indices reported; verify_batch on the same input says Verify."""
    import numpy as np
    n, nk = (1 << 17) + 9, 61
    seeds_k = np.stack([np.frombuffer(hashlib.sha512(b"e%d" % k).digest()[:32], dtype=np.uint8) for k in range(nk)])
    seeds = np.ascontiguousarray(seeds_k[np.arange(n) % nk])
    lens = (np.arange(n) % 5) * 13 + 1
    offs = np.zeros(n + 1, dtype=np.uint64); offs[1:] = np.cumsum(lens)
    fl = np.random.Generator(np.random.PCG64(8)).integers(0, 256, size=int(offs[-1]), dtype=np.uint8)
    pks, sigs = eng.sign_batch_flat(seeds, fl, offs, n)
    pk = np.frombuffer(pks, dtype=np.uint8).copy(); sg = np.frombuffer(sigs, dtype=np.uint8).copy()
    rc, res = eng.verify_each_flat(fl, offs, sg, pk, n)
    assert rc == OK and not any(res)
    bad = [0, 65535, 65536, 100000, n - 1]
    for i in bad:
        sg[64 * i + 33] ^= 2
    rc, res = eng.verify_each_flat(fl, offs, sg, pk, n, strict=True)
    assert rc == VERIFY and [i for i, r in enumerate(res) if r] == bad
    assert eng.verify_batch_flat(fl, offs, sg, pk, n) == VERIFY
    for i in bad[:2]:
        m = fl[int(offs[i]):int(offs[i + 1])].tobytes()
        assert oracle.verify(m, sg[64 * i:64 * i + 64].tobytes(), pk[32 * i:32 * i + 32].tobytes()) == res[i]
