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


@pytest.mark.parametrize("mode", [2, 0])
def test_comb_path_on_reference_fixtures(eng, oracle, mode):
    """The per-key comb tables (option each_comb: 2 = always, 0 = never, 1 = when keys repeat >= 8
    times on average) must give the verdicts of the plain kernel and of the oracle on every reference fixture: the 914
    VALIDATIONVECTORS (small-order and mixed-order keys, non-canonical encodings), the TESTVECTORS with one failure of every
    kind, from host and from device buffers."""
    import numpy as np
    import torch
    with open(os.path.join(ROOT, "tests", "golden", "ed25519_validation.json")) as f:
        vec = json.load(f)["vectors"]
    with open(os.path.join(ROOT, "tests", "golden", "ed25519_testvectors.json")) as f:
        tv = json.load(f)["vectors"]
    msgs = [v["msg"].encode() for v in vec] + [H(v["msg"]) for v in tv]
    sigs = [H(v["sig"]) for v in vec] + [H(v["sig"]) for v in tv]
    keys = [H(v["key"]) for v in vec] + [H(v["pk"]) for v in tv]
    n0 = len(vec)
    msgs[n0 + 3] += b"x"
    x = bytearray(sigs[n0 + 10]); x[63] |= 0xf0; sigs[n0 + 10] = bytes(x)
    keys[n0 + 20] = (2).to_bytes(32, "little")
    sigs[n0 + 40] = (2).to_bytes(32, "little") + sigs[n0 + 40][32:]
    # every signature three times: keys repeat
    msgs, sigs, keys = msgs * 3, sigs * 3, keys * 3
    n = len(msgs)
    fl, offs = flat(msgs)
    eng.set_option("each_comb", mode)
    try:
        for strict in (False, True):
            want = [oracle.verify(m, s, k, strict=strict) for m, s, k in zip(msgs[:n // 3], sigs[:n // 3], keys[:n // 3])] * 3
            rc, res = eng.verify_each_flat(fl, offs, b"".join(sigs), b"".join(keys), n, strict=strict)
            assert res == want and rc == VERIFY, strict
            dev = torch.device("cuda", 0)
            d = [torch.from_numpy(x_).to(dev) for x_ in (fl, offs.view(np.int64), np.frombuffer(b"".join(sigs), dtype=np.uint8).copy(),
                                                        np.frombuffer(b"".join(keys), dtype=np.uint8).copy())]
            rc, res = eng.verify_each_flat(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n, strict=strict, device_ptrs=True)
            assert res == want and rc == VERIFY, strict
    finally:
        eng.set_option("each_comb", 1)


def test_comb_path_large_repeated_keys(eng, oracle):
    """2^16 signatures by 37 keys (the automatic choice takes the comb path): all valid; planted failures are exactly the ones
    reported, with the oracle's error kinds."""
    import numpy as np
    n, nk = 1 << 16, 37
    seeds_k = np.stack([np.frombuffer(hashlib.sha512(b"c%d" % k).digest()[:32], dtype=np.uint8) for k in range(nk)])
    seeds = np.ascontiguousarray(seeds_k[np.arange(n) % nk])
    offs = np.arange(n + 1, dtype=np.uint64) * 40
    fl = np.random.Generator(np.random.PCG64(18)).integers(0, 256, size=40 * n, dtype=np.uint8)
    pks, sigs = eng.sign_batch_flat(seeds, fl, offs, n)
    pk = np.frombuffer(pks, dtype=np.uint8).copy(); sg = np.frombuffer(sigs, dtype=np.uint8).copy()
    rc, res = eng.verify_each_flat(fl, offs, sg, pk, n)
    assert rc == OK and not any(res)
    bad = [0, 777, 40000, n - 1]
    for i in bad:
        fl[40 * i + 7] ^= 0x20
    sg[64 * 1234 + 63] |= 0xf0                                       # ScalarFormat
    rc, res = eng.verify_each_flat(fl, offs, sg, pk, n)
    assert rc == VERIFY and [i for i, r in enumerate(res) if r] == sorted(bad + [1234])
    for i in (0, 1234):
        assert res[i] == oracle.verify(fl[40 * i:40 * i + 40].tobytes(), sg[64 * i:64 * i + 64].tobytes(), pk[32 * i:32 * i + 32].tobytes())
