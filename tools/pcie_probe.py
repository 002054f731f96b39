"""Host-to-device copy rates on this box (pinned memory): one large 1-D copy, the 2^20 x 160-byte point array in four
chunks as the engine issues it, and a pitched 2-D copy of 120 of every 160 bytes (X | Y | Z without T).  JSON to stdout."""
import ctypes as C
import json
import time

import torch

rt = C.CDLL("libcudart.so")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
H2D = 1
n = 1 << 20
h = torch.empty(n * 160, dtype=torch.uint8).pin_memory()
h.random_(0, 255)
d = torch.empty(n * 160, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


out = {}
t = timed(lambda: rt.cudaMemcpyAsync(d.data_ptr(), h.data_ptr(), n * 160, H2D, st))
out["1d_160MB"] = {"ms": t * 1e3, "GB_per_s": n * 160 / t / 1e9}


def four():
    for k in range(4):
        off = k * (n // 4) * 160
        rt.cudaMemcpyAsync(d.data_ptr() + off, h.data_ptr() + off, (n // 4) * 160, H2D, st)


t = timed(four)
out["1d_four_chunks"] = {"ms": t * 1e3, "GB_per_s": n * 160 / t / 1e9}
for width, dpitch in ((120, 160), (120, 120), (128, 160), (80, 160)):
    t = timed(lambda: rt.cudaMemcpy2DAsync(d.data_ptr(), dpitch, h.data_ptr(), 160, width, n, H2D, st))
    out["2d_width%d_of_160_dpitch%d" % (width, dpitch)] = {"ms": t * 1e3, "payload_GB_per_s": n * width / t / 1e9, "ms_vs_1d_of_160": None}
print(json.dumps(out, indent=1))

# ---- does the way the pinned buffer is filled matter (NUMA first touch)?  bench.py used torch.from_numpy(x).pin_memory()
import numpy as np
res = {}
src = np.random.default_rng(1).integers(0, 255, size=n * 160, dtype=np.uint8)
for label, threads in (("from_numpy.pin_memory, default threads", None), ("from_numpy.pin_memory, 1 thread", 1)):
    if threads:
        torch.set_num_threads(threads)
    hb = torch.from_numpy(src).pin_memory()
    t = timed(lambda: rt.cudaMemcpyAsync(d.data_ptr(), hb.data_ptr(), n * 160, H2D, st))
    res[label] = {"ms": t * 1e3, "GB_per_s": n * 160 / t / 1e9}
    del hb
hb = torch.empty(n * 160, dtype=torch.uint8).pin_memory()
hb.copy_(torch.from_numpy(src))
t = timed(lambda: rt.cudaMemcpyAsync(d.data_ptr(), hb.data_ptr(), n * 160, H2D, st))
res["empty.pin_memory then copy_ (1 thread)"] = {"ms": t * 1e3, "GB_per_s": n * 160 / t / 1e9}
# the bench's own buffers
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import curve25519_dalek_b200 as pkg
import bench
eng = pkg.Engine(0)
wl = bench.MsmWorkload(eng, n, n, 0, torch)
t = timed(lambda: rt.cudaMemcpyAsync(d.data_ptr(), wl.h_points.data_ptr(), n * 160, H2D, st))
res["bench.MsmWorkload.h_points"] = {"ms": t * 1e3, "GB_per_s": n * 160 / t / 1e9}
print(json.dumps(res, indent=1))
