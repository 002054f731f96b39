"""Builds the TEST-ONLY device library tests/device/libfe_device_check.so (sm_100a, in-tree so that it travels to the
GPU box).  It compiles the product's field headers into a tiny harness; the product library does not contain it."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "fe_device_check.cu")
OUT = os.path.join(HERE, "libfe_device_check.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build():
    csrc = os.path.join(ROOT, "curve25519_dalek_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("fe.cuh", "fe64.cuh")]
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    subprocess.check_call([NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                           "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC, "-lcudart"])
    return OUT


if __name__ == "__main__":
    print(build())
