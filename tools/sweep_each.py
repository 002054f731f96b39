"""verify_each 2^20 signatures (device-resident) for library variants built with different occupancy caps
(usage: sweep_each.py libdalek_b200.so libvariant_e3.so ...)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, time
sys.path.insert(0, "%s"); sys.path.insert(0, "%s/tests")
import numpy as np, torch
import curve25519_dalek_b200 as pkg
import bench
eng = pkg.Engine(0)
n = 1 << 20
flat, offs, sigs, pks = bench.build_verify_inputs(eng, n)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(x if x.dtype == np.uint8 else x.view(np.int64)).to(dev) for x in (flat, offs, sigs, pks)]
res = np.zeros(n, dtype=np.uint8)
for strict in (0, 1):
    for _ in range(2):
        assert eng.lib.ed25519_b200_verify_each_flat_dev(eng.h, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n, strict, res.ctypes.data) == 0
    t0 = time.perf_counter()
    for _ in range(4):
        eng.lib.ed25519_b200_verify_each_flat_dev(eng.h, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n, strict, res.ctypes.data)
    print("strict=%%d %%.2f ms (%%.1f M sigs/s)" %% (strict, (time.perf_counter() - t0) / 4 * 1e3, n / ((time.perf_counter() - t0) / 4) / 1e6), flush=True)
''' % (root, root)
for lib in sys.argv[1:]:
    env = dict(os.environ, DALEK_B200_LIB=os.path.join(root, "curve25519_dalek_b200", lib))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(lib, "|", r.stdout.strip().replace("\n", " | "), r.stderr.strip()[-300:], flush=True)
