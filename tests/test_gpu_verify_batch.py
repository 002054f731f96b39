"""GPU parity tests for ed25519 verify_batch: verdicts and z_i coefficients of the CUDA engine
against the CPU oracle (E/batch.rs:146-251), negative controls, chunked transcripts, fixtures."""
import json
import os
import random

import pytest

import pyref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OK, VERIFY, ARRAYLEN, SCALARFMT, POINTDEC = 0, 1, 2, 3, 4


@pytest.fixture(scope="module")
def eng():
    import curve25519_dalek_b200 as pkg
    e = pkg.Engine(0)
    yield e
    e.close()


def make_batch(oracle, n, seed=0, msg_len=None):
    rnd = random.Random(seed)
    msgs, sigs, pks = [], [], []
    for _ in range(n):
        sk = rnd.randbytes(32)
        m = rnd.randbytes(msg_len if msg_len is not None else rnd.randrange(0, 300))
        msgs.append(m); pks.append(oracle.public_key(sk)); sigs.append(oracle.sign(m, sk))
    return msgs, sigs, pks


CHUNK = 0           # the engine's default verify_chunk: 0 = the reference's single transcript over the whole batch


def run(eng, msgs, sigs, pks):
    return eng.verify_batch_raw(msgs, b"".join(sigs), b"".join(pks))


@pytest.mark.parametrize("n", [0, 1, 2, 7, 63, 64, 65, 95, 96, 128, 129, 300])
def test_verify_batch_valid_and_zs(eng, oracle, n):
    msgs, sigs, pks = make_batch(oracle, n, seed=n)
    chunk = CHUNK
    rc_o, zs_o = oracle.verify_batch(msgs, sigs, pks, chunk=chunk, want_zs=True)
    assert rc_o == OK
    assert run(eng, msgs, sigs, pks) == OK
    assert eng.last_zs(n) == zs_o            # every z_i is the reference's (one transcript over the whole batch)


def test_verify_batch_chunk_option(eng, oracle):
    msgs, sigs, pks = make_batch(oracle, 50, seed=77, msg_len=59)
    for chunk in (1, 3, 16, 50, 4096):
        eng.set_option("verify_chunk", chunk)
        try:
            assert run(eng, msgs, sigs, pks) == OK
            rc, zs = oracle.verify_batch(msgs, sigs, pks, chunk=chunk, want_zs=True)
            assert eng.last_zs(50) == zs
        finally:
            eng.set_option("verify_chunk", 0)


def test_transcript_kernels_agree(eng, oracle):
    """The Merlin transcript runs on one WARP (25-lane Keccak, csrc/transcript_warp.cuh) when a launch has few transcripts
    and on one thread per transcript otherwise; option `transcript_warp` selects.  Both must draw the oracle's z_i: one
    transcript over 301 signatures (n = 53 and 219 end exactly on a rate-block boundary), and chunks of 7."""
    for n in (1, 2, 53, 219, 301):
        msgs, sigs, pks = make_batch(oracle, n, seed=5000 + n, msg_len=33)
        for chunk in (0, 7):
            rc_o, zs_o = oracle.verify_batch(msgs, sigs, pks, chunk=chunk, want_zs=True)
            assert rc_o == OK
            # warp = 1: one warp per transcript; warp = 0: one thread each, with the rate block staged in shared memory
            # (blocks = 1, k_transcript_blocks) or byte by byte on the sponge state (blocks = 0, k_transcript)
            for warp, blocks in ((1, 1), (0, 1), (0, 0)):
                eng.set_option("transcript_warp", warp)
                eng.set_option("transcript_blocks", blocks)
                eng.set_option("verify_chunk", chunk)
                try:
                    assert run(eng, msgs, sigs, pks) == OK
                    assert eng.last_zs(n) == zs_o, (n, chunk, warp, blocks)
                finally:
                    eng.set_option("transcript_warp", 1)
                    eng.set_option("transcript_blocks", 1)
                    eng.set_option("verify_chunk", 0)


def test_verify_batch_negative_controls(eng, oracle):
    msgs, sigs, pks = make_batch(oracle, 33, seed=99, msg_len=59)
    cases = []
    bad = list(sigs); b = bytearray(bad[4]); b[33] ^= 1; bad[4] = bytes(b)
    cases.append((msgs, bad, pks))                                   # bit flip in s
    bad = list(sigs); b = bytearray(bad[9]); b[3] ^= 0x10; bad[9] = bytes(b)
    cases.append((msgs, bad, pks))                                   # bit flip in R
    bm = list(msgs); bm[2] = bytes([bm[2][0] ^ 1]) + bm[2][1:]
    cases.append((bm, sigs, pks))                                    # bit flip in a message
    bad = list(sigs); bad[0] = (2).to_bytes(32, "little") + bad[0][32:]
    cases.append((msgs, bad, pks))                                   # R not on the curve
    s = int.from_bytes(sigs[1][32:], "little") + pyref.L
    bad = list(sigs); bad[1] = sigs[1][:32] + s.to_bytes(32, "little")
    cases.append((msgs, bad, pks))                                   # non-canonical s
    bk = list(pks); bk[3] = (2).to_bytes(32, "little")
    cases.append((msgs, sigs, bk))                                   # undecodable key
    sw = list(sigs); sw[0], sw[1] = sw[1], sw[0]
    cases.append((msgs, sw, pks))                                    # swapped signatures
    both = list(sigs); both[1] = sigs[1][:32] + s.to_bytes(32, "little"); both[0] = (2).to_bytes(32, "little") + both[0][32:]
    cases.append((msgs, both, pks))                                  # bad R and bad s: ScalarFormat wins
    want = [VERIFY, VERIFY, VERIFY, VERIFY, SCALARFMT, POINTDEC, VERIFY, SCALARFMT]
    for (m, s_, k), w in zip(cases, want):
        assert oracle.verify_batch(m, s_, k) == w
        assert run(eng, m, s_, k) == w


def test_verify_batch_reference_fixtures(eng, oracle):
    """TESTVECTORS (valid signatures) verify as one batch; every VALIDATIONVECTORS case gets the
    same verdict from the engine as from the oracle's verify_batch on that single signature."""
    with open(os.path.join(ROOT, "tests", "golden", "ed25519_testvectors.json")) as f:
        tv = json.load(f)["vectors"]
    H = bytes.fromhex
    msgs, sigs, pks = [H(v["msg"]) for v in tv], [H(v["sig"]) for v in tv], [H(v["pk"]) for v in tv]
    assert run(eng, msgs, sigs, pks) == OK
    with open(os.path.join(ROOT, "tests", "golden", "ed25519_validation.json")) as f:
        vv = json.load(f)["vectors"]
    assert len(vv) == 914
    for v in vv:                                                     # all 914 cases, each as a batch of one
        m, s_, k = [v["msg"].encode()], [H(v["sig"])], [H(v["key"])]
        assert run(eng, m, s_, k) == oracle.verify_batch(m, s_, k), v["number"]
    # ... and all of them in ONE call as 914 independent batches of one signature
    msgs = [v["msg"].encode() for v in vv]
    flat, offs = _flat(msgs)
    rc, verdicts = eng.verify_batches_flat(flat, offs, b"".join(H(v["sig"]) for v in vv), b"".join(H(v["key"]) for v in vv), len(vv), 1)
    assert verdicts == [oracle.verify_batch([m], [H(v["sig"])], [H(v["key"])]) for m, v in zip(msgs, vv)]
    # ... and as batches of 3 and of 32 (ragged tail): the small-order defects of different batches cancel in sums of
    # batch equations one time in eight, so every batch's own small-order part has to be tested (k_batch_torsion)
    sg, ks = [H(v["sig"]) for v in vv], [H(v["key"]) for v in vv]
    for bs in (3, 32):
        for dedupe in (1, 0):
            eng.set_option("dedupe_keys", dedupe)
            try:
                rc, verdicts = eng.verify_batches_flat(flat, offs, b"".join(sg), b"".join(ks), len(vv), bs)
            finally:
                eng.set_option("dedupe_keys", 1)
            assert verdicts == [oracle.verify_batch(msgs[k:k + bs], sg[k:k + bs], ks[k:k + bs]) for k in range(0, len(vv), bs)], (bs, dedupe)


def test_verify_batch_python_api(eng, oracle):
    import curve25519_dalek_b200 as pkg
    msgs, sigs, pks = make_batch(oracle, 12, seed=5)
    assert pkg.verify_batch(msgs, sigs, pks, engine=eng) is None
    with pytest.raises(pkg.SignatureError) as ei:
        pkg.verify_batch(msgs, sigs[:-1], pks, engine=eng)
    assert ei.value.kind == "ArrayLength"
    bad = list(sigs); b = bytearray(bad[4]); b[40] ^= 1; bad[4] = bytes(b)
    with pytest.raises(pkg.SignatureError) as ei:
        pkg.verify_batch(msgs, bad, pks, engine=eng)
    assert ei.value.kind == "Verify"


def test_verify_batch_flat_large(eng, oracle):
    """n = 20000 with 64 distinct keys: all valid -> Ok; one corrupted message -> Verify."""
    import numpy as np
    n, nk = 20000, 64
    rnd = random.Random(4242)
    seeds = [rnd.randbytes(32) for _ in range(nk)]
    keys = [oracle.public_key(s) for s in seeds]
    msgs = [(b"a" * 51) + i.to_bytes(8, "little") for i in range(n)]
    sigs = [oracle.sign(msgs[i], seeds[i % nk]) for i in range(n)]
    flat = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy()
    offs = np.arange(n + 1, dtype=np.uint64) * 59
    sg = np.frombuffer(b"".join(sigs), dtype=np.uint8).copy()
    pk = np.frombuffer(b"".join(keys[i % nk] for i in range(n)), dtype=np.uint8).copy()
    assert eng.verify_batch_flat(flat, offs, sg, pk, n) == OK
    rc_o, zs_o = oracle.verify_batch(msgs, sigs, [keys[i % nk] for i in range(n)], want_zs=True)
    assert rc_o == OK and eng.last_zs(n) == zs_o             # 20000 coefficients of the reference's single transcript
    flat[59 * 12345 + 7] ^= 1
    assert eng.verify_batch_flat(flat, offs, sg, pk, n) == VERIFY


def test_verify_batch_host_streaming_pieces(eng, oracle):
    """n = 2^18 + 77 through the host-buffer entry point: the batch is streamed in pieces (copy of piece
    k+1 overlaps hashing / decompression of piece k).  Signatures come from the GPU signer (byte-exact
    on TESTVECTORS, see test_gpu_straus_base.py); a sample is re-verified by the oracle."""
    import hashlib
    import numpy as np
    n, nk = (1 << 18) + 77, 128
    seeds_k = np.stack([np.frombuffer(hashlib.sha512(b"k%d" % k).digest()[:32], dtype=np.uint8) for k in range(nk)])
    seeds = np.ascontiguousarray(seeds_k[np.arange(n) % nk])
    lens = (np.arange(n) % 7) * 9 + 3                       # ragged message lengths 3..57
    offs = np.zeros(n + 1, dtype=np.uint64); offs[1:] = np.cumsum(lens)
    rng = np.random.Generator(np.random.PCG64(5))
    flat = rng.integers(0, 256, size=int(offs[-1]), dtype=np.uint8)
    pks, sigs = eng.sign_batch_flat(seeds, flat, offs, n)
    pk = np.frombuffer(pks, dtype=np.uint8).copy(); sg = np.frombuffer(sigs, dtype=np.uint8).copy()
    for i in (0, 1, 70000, n - 1):                          # oracle spot checks of the synthesised inputs
        m = flat[int(offs[i]):int(offs[i + 1])].tobytes()
        assert oracle.verify(m, sg[64 * i:64 * i + 64].tobytes(), pk[32 * i:32 * i + 32].tobytes()) == OK
    eng.set_option("verify_chunk", 64)                      # opt-in chunked transcripts: the piece sweep runs 6 calls
    for pieces in (4, 1, 3):
        eng.set_option("verify_pieces", pieces)
        try:
            assert eng.verify_batch_flat(flat, offs, sg, pk, n) == OK
            zs = eng.last_zs(n)
            sg[64 * (n - 5) + 40] ^= 1                      # corrupt s of a signature in the last piece
            assert eng.verify_batch_flat(flat, offs, sg, pk, n) == VERIFY
            sg[64 * (n - 5) + 40] ^= 1
        finally:
            eng.set_option("verify_pieces", 4)
        if pieces == 4:
            zs4 = zs
        else:
            assert zs == zs4                                # the coefficients do not depend on the piece count
    # z_i of the first transcript chunk agree with the oracle's
    first = 64
    msgs = [flat[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(first)]
    rc, zo = oracle.verify_batch(msgs, [sg[64 * i:64 * i + 64].tobytes() for i in range(first)],
                                 [pk[32 * i:32 * i + 32].tobytes() for i in range(first)], chunk=64, want_zs=True)
    assert rc == OK and zs4[:16 * first] == zo
    eng.set_option("verify_chunk", 0)
    # default mode (one transcript over all 2^18 + 77 signatures, streamed in 4 pieces): verdict only
    assert eng.verify_batch_flat(flat, offs, sg, pk, n) == OK


def test_verify_batch_key_dedupe(eng, oracle):
    """Repeated public keys are decompressed once (dedupe_keys option): verdicts, coefficients and error
    codes must not depend on the option, including an undecodable key that appears several times."""
    rnd = random.Random(77)
    seeds = [rnd.randbytes(32) for _ in range(5)]
    msgs = [rnd.randbytes(40) for _ in range(60)]
    pks = [oracle.public_key(seeds[i % 5]) for i in range(60)]
    sigs = [oracle.sign(msgs[i], seeds[i % 5]) for i in range(60)]
    bad_key = (2).to_bytes(32, "little")
    for opt in (1, 0):
        eng.set_option("dedupe_keys", opt)
        try:
            assert run(eng, msgs, sigs, pks) == OK
            zs = eng.last_zs(60)
            b = list(sigs); x = bytearray(b[17]); x[35] ^= 4; b[17] = bytes(x)
            assert run(eng, msgs, b, pks) == VERIFY == oracle.verify_batch(msgs, b, pks)
            k = list(pks); k[3] = bad_key; k[44] = bad_key
            assert run(eng, msgs, sigs, k) == POINTDEC == oracle.verify_batch(msgs, sigs, k)
            wrong = list(pks); wrong[10] = pks[11]                 # valid key, wrong signer
            assert run(eng, msgs, sigs, wrong) == VERIFY
        finally:
            eng.set_option("dedupe_keys", 1)
        if opt == 1:
            z1 = zs
        else:
            assert zs == z1


def test_verify_batch_key_merging_edges(eng, oracle):
    """One MSM term per distinct key (scalar = sum of z_i h_i over its signatures): n-1 distinct keys (a single
    merged pair), every signature under one key, and a swap of two signatures of the same key (each valid for the
    other's message only) must all give the oracle's verdict."""
    msgs, sigs, pks = make_batch(oracle, 150, seed=901)
    rnd = random.Random(902)
    sk = rnd.randbytes(32)
    msgs[77] = b"second message of key 3"; pks[77] = pks[3]
    # signature 77 must really be by key 3: rebuild both from one seed
    pks[3] = pks[77] = oracle.public_key(sk)
    sigs[3], sigs[77] = oracle.sign(msgs[3], sk), oracle.sign(msgs[77], sk)
    assert run(eng, msgs, sigs, pks) == OK == oracle.verify_batch(msgs, sigs, pks, chunk=CHUNK)
    swapped = list(sigs); swapped[3], swapped[77] = sigs[77], sigs[3]
    assert run(eng, msgs, swapped, pks) == VERIFY == oracle.verify_batch(msgs, swapped, pks, chunk=CHUNK)
    one_key = [oracle.public_key(sk)] * 150
    one_sigs = [oracle.sign(m, sk) for m in msgs]
    assert run(eng, msgs, one_sigs, one_key) == OK
    bad = list(one_sigs); x = bytearray(bad[149]); x[2] ^= 0x10; bad[149] = bytes(x)      # R of the last signature
    assert run(eng, msgs, bad, one_key) == VERIFY == oracle.verify_batch(msgs, bad, one_key, chunk=CHUNK)


def _flat(msgs):
    import numpy as np
    offs = np.zeros(len(msgs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(m) for m in msgs])
    return np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy(), offs


@pytest.mark.parametrize("dedupe", [1, 0])
def test_verify_batches_independent_verdicts(eng, oracle, dedupe):
    """Many independent batches in one call: verdicts[k] must be what the oracle's
    verify_batch returns for batch k alone, the z_i those of one reference transcript per batch, for clean input,
    for every error kind, and for failures in first / middle / last (ragged) batches."""
    bs, n = 16, 630                                            # 40 batches, the last one has 6 signatures
    rnd = random.Random(1234)
    seeds = [rnd.randbytes(32) for _ in range(9)]
    msgs = [rnd.randbytes(rnd.randrange(0, 90)) for _ in range(n)]
    pks = [oracle.public_key(seeds[i % 9]) for i in range(n)]
    sigs = [oracle.sign(msgs[i], seeds[i % 9]) for i in range(n)]
    nb = (n + bs - 1) // bs

    def expect(m, s, k):
        out = []
        for b in range(nb):
            lo, hi = b * bs, min(n, (b + 1) * bs)
            out.append(oracle.verify_batch(m[lo:hi], s[lo:hi], k[lo:hi]))
        return out

    def got(m, s, k):
        flat, offs = _flat(m)
        return eng.verify_batches_flat(flat, offs, b"".join(s), b"".join(k), n, bs)

    eng.set_option("dedupe_keys", dedupe)
    try:
        rc, v = got(msgs, sigs, pks)
        assert rc == OK and v == [OK] * nb == expect(msgs, sigs, pks)
        zs = eng.last_zs(n)
        for b in (0, 17, nb - 1):                              # one reference transcript per batch
            lo, hi = b * bs, min(n, (b + 1) * bs)
            rc_o, z_o = oracle.verify_batch(msgs[lo:hi], sigs[lo:hi], pks[lo:hi], want_zs=True)
            assert rc_o == OK and zs[16 * lo:16 * hi] == z_o
        # failures of every kind, spread over the call
        m, s, k = list(msgs), list(sigs), list(pks)
        m[5] = m[5] + b"!"                                     # batch 0: Verify
        x = bytearray(s[16 * 7 + 3]); x[63] |= 0xf0; s[16 * 7 + 3] = bytes(x)        # batch 7: s >= l -> ScalarFormat
        k[16 * 11 + 9] = (2).to_bytes(32, "little")            # batch 11: undecodable key -> PointDecompression
        s[16 * 20] = (2).to_bytes(32, "little") + s[16 * 20][32:]                  # batch 20: undecodable R -> Verify
        x = bytearray(s[16 * 25 + 1]); x[63] |= 0xf0; s[16 * 25 + 1] = bytes(x)      # batch 25: bad s AND bad key -> PointDecompression
        k[16 * 25 + 2] = (2).to_bytes(32, "little")
        x = bytearray(s[n - 1]); x[40] ^= 1; s[n - 1] = bytes(x)                   # last (ragged) batch: Verify
        want = expect(m, s, k)
        assert want[0] == VERIFY and want[7] == SCALARFMT and want[11] == POINTDEC and want[20] == VERIFY and want[25] == POINTDEC
        assert want[nb - 1] == VERIFY and sum(1 for w in want if w) == 6
        rc, v = got(m, s, k)
        assert rc == VERIFY and v == want
        # a single bad signature in the middle: exactly one verdict changes
        m2 = list(msgs); m2[16 * 19 + 15] = b"x" + m2[16 * 19 + 15]
        rc, v = got(m2, sigs, pks)
        assert rc == VERIFY and v == [VERIFY if b == 19 else OK for b in range(nb)]
    finally:
        eng.set_option("dedupe_keys", 1)


def test_verify_batches_large_device(eng, oracle):
    """2^15 signatures as 128 batches of 256 (the reference's largest published batch size) from device memory."""
    import hashlib
    import numpy as np
    import torch
    n, bs, nk = 1 << 15, 256, 37
    seeds_k = np.stack([np.frombuffer(hashlib.sha512(b"b%d" % k).digest()[:32], dtype=np.uint8) for k in range(nk)])
    seeds = np.ascontiguousarray(seeds_k[np.arange(n) % nk])
    offs = np.arange(n + 1, dtype=np.uint64) * 59
    flat = np.random.Generator(np.random.PCG64(6)).integers(0, 256, size=59 * n, dtype=np.uint8)
    pks, sigs = eng.sign_batch_flat(seeds, flat, offs, n)
    pk = np.frombuffer(pks, dtype=np.uint8).copy(); sg = np.frombuffer(sigs, dtype=np.uint8).copy()
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(x if x.dtype == np.uint8 else x.view(np.int64)).to(dev) for x in (flat, offs, sg, pk)]
    rc, v = eng.verify_batches_flat(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n, bs, device_ptrs=True)
    assert rc == OK and v == [OK] * (n // bs)
    zs = eng.last_zs(n)
    b = 77                                                     # one batch against the oracle's single transcript
    lo, hi = b * bs, (b + 1) * bs
    msgs = [flat[59 * i:59 * i + 59].tobytes() for i in range(lo, hi)]
    rc_o, z_o = oracle.verify_batch(msgs, [sg[64 * i:64 * i + 64].tobytes() for i in range(lo, hi)],
                                    [pk[32 * i:32 * i + 32].tobytes() for i in range(lo, hi)], want_zs=True)
    assert rc_o == OK and zs[16 * lo:16 * hi] == z_o
    d[0][59 * (bs * 100 + 3) + 1] ^= 4                         # corrupt one message of batch 100 and one of batch 0
    d[0][7] ^= 1
    rc, v = eng.verify_batches_flat(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n, bs, device_ptrs=True)
    assert rc == VERIFY and v == [VERIFY if k in (0, 100) else OK for k in range(n // bs)]
    # the option set for the call is restored afterwards
    assert eng.verify_batch_flat(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n, device_ptrs=True) == VERIFY


@pytest.mark.parametrize("dedupe", [1, 0])
def test_verify_batches_host_pieces(eng, oracle, dedupe):
    """Host buffers of >= 2^18 signatures are streamed in pieces (copies of piece k+1 under the front end of piece k; keys are
    de-duplicated across pieces): failures planted in different pieces, a key that only occurs late, a ragged last
    batch; the same verdicts from device memory and for every piece count; three planted batches against the oracle's
    verify_batch."""
    import hashlib
    import numpy as np
    import torch
    n, bs, nk = (1 << 18) + 300, 256, 61
    seeds_k = np.stack([np.frombuffer(hashlib.sha512(b"p%d" % k).digest()[:32], dtype=np.uint8) for k in range(nk + 1)])
    idx = np.arange(n) % nk
    idx[n - 5000:] = np.where(np.arange(5000) % 7 == 0, nk, idx[n - 5000:])       # key nk only occurs in the last piece
    seeds = np.ascontiguousarray(seeds_k[idx])
    offs = np.arange(n + 1, dtype=np.uint64) * 59
    flat = np.random.Generator(np.random.PCG64(9)).integers(0, 256, size=59 * n, dtype=np.uint8)
    pks, sigs = eng.sign_batch_flat(seeds, flat, offs, n)
    pk = np.frombuffer(pks, dtype=np.uint8).copy(); sg = np.frombuffer(sigs, dtype=np.uint8).copy()
    nb = (n + bs - 1) // bs
    bad = {3: 3 * bs + 17, 400: 400 * bs + 255, 777: 777 * bs, nb - 1: n - 1}      # batch -> corrupted signature
    for i in bad.values():
        flat[59 * i + 5] ^= 0x10
    sg[64 * (600 * bs + 9) + 63] |= 0x80                                           # batch 600: non-canonical s -> ScalarFormat
    want = [OK] * nb
    for k in bad:
        want[k] = VERIFY
    want[600] = SCALARFMT
    eng.set_option("dedupe_keys", dedupe)
    try:
        for pieces in (4, 3, 8, 1):
            eng.set_option("verify_pieces", pieces)
            rc, v = eng.verify_batches_flat(flat, offs, sg.tobytes(), pk.tobytes(), n, bs)
            assert rc == VERIFY and v == want, pieces
        dev = torch.device("cuda", 0)
        d = [torch.from_numpy(x if x.dtype == np.uint8 else x.view(np.int64)).to(dev) for x in (flat, offs, sg, pk)]
        rc, v = eng.verify_batches_flat(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n, bs, device_ptrs=True)
        assert rc == VERIFY and v == want
    finally:
        eng.set_option("verify_pieces", 4)
        eng.set_option("dedupe_keys", 1)
    for k in (3, 600, nb - 1):
        lo, hi = k * bs, min(n, (k + 1) * bs)
        assert oracle.verify_batch([flat[59 * i:59 * i + 59].tobytes() for i in range(lo, hi)], [sg[64 * i:64 * i + 64].tobytes() for i in range(lo, hi)],
                                   [pk[32 * i:32 * i + 32].tobytes() for i in range(lo, hi)]) == want[k]


def test_verify_batches_all_distinct_keys_with_failure(eng, oracle):
    """Every key different (no merging possible) and one bad batch: the bisection evaluates sub-ranges through the
    per-key accumulation path; exactly that batch must fail, with and without key merging."""
    msgs, sigs, pks = make_batch(oracle, 100, seed=4321)
    bs, nb = 16, 7
    flat, offs = _flat(msgs)
    for dedupe in (1, 0):
        eng.set_option("dedupe_keys", dedupe)
        try:
            rc, v = eng.verify_batches_flat(flat, offs, b"".join(sigs), b"".join(pks), 100, bs)
            assert rc == OK and v == [OK] * nb
            bad = list(sigs); x = bytearray(bad[16 * 4 + 9]); x[33] ^= 8; bad[16 * 4 + 9] = bytes(x)
            rc, v = eng.verify_batches_flat(flat, offs, b"".join(bad), b"".join(pks), 100, bs)
            assert rc == VERIFY and v == [VERIFY if k == 4 else OK for k in range(nb)]
            want = [oracle.verify_batch(msgs[k * bs:(k + 1) * bs], bad[k * bs:(k + 1) * bs], pks[k * bs:(k + 1) * bs]) for k in range(nb)]
            assert v == want
        finally:
            eng.set_option("dedupe_keys", 1)


# ---- small-order components: the verdict of the un-cofactored batch equation depends on the z_i (batch.rs:240-250) ----
@pytest.mark.parametrize("n", [5, 100, 300])
def test_verify_batch_mixed_order_components_match_reference(eng, oracle, n):
    """Signatures whose R, or whose public key A, carries a small-order component leave a pure torsion defect in the
    batch equation; Ok or Verify then depends on the z_i modulo 8 (R) and on (z_i h_i mod l) modulo 8 (A), so only the
    reference's own coefficients -- ONE transcript over the whole batch, also for n > 64 -- and exact per-key scalar
    sums (modulo 8 l, not l) give the reference's verdict.  tests/test_oracle_ed25519.py shows on the CPU that both
    verdicts occur over these seeds and that chunked transcripts would disagree."""
    import torsion_cases
    outcomes = set()
    for trial in range(10):
        msgs, sigs, pks = torsion_cases.make_batch(oracle, n, seed=1000 * n + trial)
        want = oracle.verify_batch(msgs, sigs, pks)
        for dedupe in (1, 0):
            eng.set_option("dedupe_keys", dedupe)
            try:
                assert run(eng, msgs, sigs, pks) == want, (trial, dedupe)
            finally:
                eng.set_option("dedupe_keys", 1)
        # the same signatures as independent batches of 16: each verdict is the reference's for that batch alone
        flat, offs = _flat(msgs)
        rc, v = eng.verify_batches_flat(flat, offs, b"".join(sigs), b"".join(pks), n, 16)
        assert v == [oracle.verify_batch(msgs[k:k + 16], sigs[k:k + 16], pks[k:k + 16]) for k in range(0, n, 16)], trial
        outcomes.add(want)
    assert outcomes <= {OK, VERIFY}


def test_verify_batch_with_key_points(eng, oracle):
    """ed25519_b200_verify_batch_flat_points: the caller passes the decompressed point of every key (what the reference's
    VerifyingKey holds, verifying.rs:65-71); verdicts and coefficients equal the byte-keyed call and the oracle, with Z = 1
    points, with projectively scaled points (Z != 1), with repeated keys and with all keys distinct."""
    import ctypes as C
    import numpy as np
    for n, distinct in ((60, False), (130, True)):
        if distinct:
            msgs, sigs, pks = make_batch(oracle, n, seed=31)
        else:
            rnd = random.Random(7)
            seeds = [rnd.randbytes(32) for _ in range(5)]
            msgs = [rnd.randbytes(40) for _ in range(n)]
            pks = [oracle.public_key(seeds[i % 5]) for i in range(n)]
            sigs = [oracle.sign(msgs[i], seeds[i % 5]) for i in range(n)]
        flat, offs = _flat(msgs)
        sg, pk = b"".join(sigs), b"".join(pks)
        rc, limbs, ok = eng.decompress_batch(pk, n)
        assert rc == 0 and all(ok)
        pts = np.frombuffer(limbs, dtype=np.uint64).copy()
        # every third point projectively rescaled (Z != 1): same point, different limbs
        for i in range(0, n, 3):
            p = oracle.p3_from_limbs([int(x) for x in pts[20 * i:20 * i + 20]])
            q = oracle.sub(oracle.add(oracle.double(p), p), oracle.double(p))
            pts[20 * i:20 * i + 20] = oracle.p3_limbs(q)
        rc_o, zs_o = oracle.verify_batch(msgs, sigs, pks, want_zs=True)
        assert rc_o == OK
        assert eng.verify_batch_flat_points(flat, offs, sg, pk, pts, n) == OK
        assert eng.last_zs(n) == zs_o
        bad = bytearray(sg); bad[64 * 17 + 40] ^= 2
        assert eng.verify_batch_flat_points(flat, offs, bytes(bad), pk, pts, n) == VERIFY
        # the point is what enters the equation: a wrong point for one key fails the batch although its bytes are right
        # (with repeated keys any signature of a key may lend it its point -- the contract is that all of them carry the
        # decoding of the key's bytes -- so this is checked where every key occurs once)
        if distinct:
            wrong = pts.copy(); wrong[20 * 4:20 * 5] = pts[20 * 5:20 * 6]
            assert eng.verify_batch_flat_points(flat, offs, sg, pk, wrong, n) == VERIFY
        # device-resident form
        import torch
        dev = torch.device("cuda", 0)
        d = [torch.from_numpy(x if x.dtype == np.uint8 else x.view(np.int64)).to(dev)
             for x in (flat, offs, np.frombuffer(sg, dtype=np.uint8).copy(), np.frombuffer(pk, dtype=np.uint8).copy(), pts)]
        assert eng.verify_batch_flat_points(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(), n,
                                            device_ptrs=True) == OK
        # independent batches of 16 with the callers' points: each verdict is the reference's for that batch alone
        want = [oracle.verify_batch(msgs[k:k + 16], [bytes(bad[64 * i:64 * i + 64]) for i in range(k, min(n, k + 16))], pks[k:k + 16])
                for k in range(0, n, 16)]
        assert VERIFY in want and OK in want
        assert eng.verify_batches_flat_points(flat, offs, bytes(bad), pk, pts, n, 16) == (1, want)
        dbad = torch.from_numpy(np.frombuffer(bytes(bad), dtype=np.uint8).copy()).to(dev)
        assert eng.verify_batches_flat_points(d[0].data_ptr(), d[1].data_ptr(), dbad.data_ptr(), d[3].data_ptr(), d[4].data_ptr(), n, 16,
                                              device_ptrs=True) == (1, want)
