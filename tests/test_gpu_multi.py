"""Single-process multi-GPU MSM (dalek_b200_init_multi / dalek_b200_edwards_vartime_msm_multi): needs >= 2 GPUs, so it is
skipped on the one-GPU box of the round-end run; run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    import torch
    import curve25519_dalek_b200 as pkg
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    ndev = min(4, torch.cuda.device_count())
    m = pkg.MultiEngine(list(range(ndev)))
    e = pkg.Engine(0)
    yield m, e, ndev
    m.close(); e.close()


def _inputs(eng, n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    t = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); t[:, 31] &= 0x0f
    s = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); s[:, 31] &= 0x0f
    limbs, comp = eng.mul_base_batch(t, n)
    return s, np.frombuffer(limbs, dtype=np.uint64).copy(), np.frombuffer(comp, dtype=np.uint8).copy()


@pytest.mark.parametrize("n", [0, 5, 40000, (1 << 18) + 3])
def test_multi_msm_equals_single_device(engines, oracle, n):
    m, e, ndev = engines
    s, limbs, comp = _inputs(e, max(n, 1), seed=n + 1)
    for fmt, pts in ((1, limbs), (0, comp)):
        rc1, want, _ = e.edwards_vartime_msm(s, pts, n, point_fmt=fmt)
        rc2, got, _ = m.edwards_vartime_msm(s, pts, n, point_fmt=fmt)
        assert rc1 == rc2 == 0 and got == want, (n, fmt)
    if 0 < n <= 40000:      # the oracle on the same inputs (reference Pippenger)
        pts = [oracle.decompress(comp[32 * i:32 * i + 32].tobytes()) for i in range(n)]
        assert got == oracle.compress(oracle.msm("optional", [s[i].tobytes() for i in range(n)], pts))


def test_multi_msm_none_on_undecodable_point(engines):
    m, e, ndev = engines
    n = 1 << 17
    s, limbs, comp = _inputs(e, n, seed=9)
    comp[32 * (n - 3):32 * (n - 3) + 32] = np.frombuffer((2).to_bytes(32, "little"), dtype=np.uint8)    # y = 2 is off the curve: last shard
    rc, _, _ = m.edwards_vartime_msm(s, comp, n, point_fmt=0)
    assert rc == 1
