import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, curve25519_dalek_b200 as pkg, argparse
eng=pkg.Engine(0)
args=argparse.Namespace(steps=3, warmup=1, sigs_per_gpu=0, verify_batch_size=256, transcript_chunk=64)
v=bench.run_verify(args,0,1,0,eng=eng,steps=3,warmup=1,each=True)
print("comb", v["value"]/1e6, v["ms_per_step"], v["e2e"]["value"]/1e6)
