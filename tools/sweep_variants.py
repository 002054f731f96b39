import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, time
sys.path.insert(0, "%s"); sys.path.insert(0, "%s/tests")
import numpy as np, torch
import curve25519_dalek_b200 as pkg
import bench
eng = pkg.Engine(0)
wl = bench.MsmWorkload(eng, 1 << 20, 1 << 20, 0, torch)
for _ in range(3): wl.step_device_single()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): wl.step_device_single()
torch.cuda.synchronize(); print("msm ms/step %%.3f" %% ((time.perf_counter() - t0) / 20 * 1e3), flush=True)
''' % (root, root)
for lib in sys.argv[1:]:
    env = dict(os.environ, DALEK_B200_LIB=os.path.join(root, "curve25519_dalek_b200", lib))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(lib, r.stdout.strip(), r.stderr.strip()[-200:])
