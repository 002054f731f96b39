"""A/B of the bucket kernel's point gather: 6-8 cp.async (LDGSTS) per point against ONE TMA bulk copy
(cp.async.bulk + per-thread mbarrier, UBLKCP in SASS) per point.  Prints one JSON object; run on the B200.

    python tools/ab_tma.py > gpurun_out/ab_tma.json
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import curve25519_dalek_b200 as pkg
import bench

eng = pkg.Engine(0)
out = {"what": "k_bucket_accumulate, gather of the next point: cp.async x NQ (acc_tma=0) vs one cp.async.bulk + mbarrier (acc_tma=1)",
       "timing": "CUDA events around the bucket kernel (dalek_b200_last_kernel_ms), mean of 30 calls after 5 warm-up calls; ms_per_step = wall clock of the blocking call"}
for log2n, fmt in ((20, "extended"), (21, "extended"), (20, "compressed")):
    n = 1 << log2n
    wl = bench.MsmWorkload(eng, n, n, 0, torch)
    if fmt == "compressed":
        import numpy as np
        enc = torch.empty(32 * n, dtype=torch.uint8).pin_memory()
        assert eng.lib.dalek_b200_edwards_compress_batch(eng.h, wl.h_points.data_ptr(), n, enc.data_ptr()) == 0
        d_enc = enc.cuda()
    res = {}
    want = None
    for tma in (0, 1, 0, 1):
        eng.set_option("acc_tma", tma)

        def step():
            if fmt == "compressed":
                rc, comp, _ = eng.edwards_vartime_msm(wl.d_scalars.data_ptr(), d_enc.data_ptr(), n, point_fmt=0, device_ptrs=True)
                assert rc == 0
                return comp
            return wl.step_device_single()
        for _ in range(5):
            got = step()
        want = want or got
        assert got == want, "TMA and cp.async variants disagree"
        torch.cuda.synchronize(); t0 = time.perf_counter(); km = []
        for _ in range(30):
            step(); km.append(eng.last_kernel_ms()[0])
        torch.cuda.synchronize()
        res.setdefault("acc_tma=%d" % tma, []).append({"bucket_kernel_ms": statistics.mean(km), "ms_per_step": (time.perf_counter() - t0) / 30 * 1e3})
    out["2^%d pairs, %s points" % (log2n, fmt)] = res
    del wl
eng.set_option("acc_tma", 0)
print(json.dumps(out, indent=1))
