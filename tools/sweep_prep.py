"""verify_batch 2^22 with the decompression exponentiation on the integer / FP64 field, for library variants built
with different occupancy caps of k_prep_R (usage: sweep_prep.py libdalek_b200.so libvariant_p4.so ...)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, time
sys.path.insert(0, "%s"); sys.path.insert(0, "%s/tests")
import numpy as np, torch
import curve25519_dalek_b200 as pkg
import bench
eng = pkg.Engine(0)
n = 1 << 22
flat, offs, sigs, pks = bench.build_verify_inputs(eng, n)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(x if x.dtype == np.uint8 else x.view(np.int64)).to(dev) for x in (flat, offs, sigs, pks)]
for f64 in (0, 1, 0, 1):
    eng.set_option("decompress_f64", f64)
    for _ in range(2):
        assert eng.verify_batch_flat(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n, device_ptrs=True) == 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8):
        eng.verify_batch_flat(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n, device_ptrs=True)
    torch.cuda.synchronize()
    print("decompress_f64=%%d verify ms/step %%.2f" %% (f64, (time.perf_counter() - t0) / 8 * 1e3), flush=True)
''' % (root, root)
for lib in sys.argv[1:]:
    env = dict(os.environ, DALEK_B200_LIB=os.path.join(root, "curve25519_dalek_b200", lib))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(lib, "|", r.stdout.strip().replace("\n", " | "), r.stderr.strip()[-300:], flush=True)
