"""CPU-only checks of the boundary: the C-ABI library loads, exports every symbol that
include/dalek_b200.h declares, refuses to run without a GPU (no fallback), and the product
package never imports the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "dalek_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:dalek_b200|ed25519_b200)_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    from curve25519_dalek_b200 import build
    path = build.build()
    return C.CDLL(path)


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s


def test_init_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    lib.dalek_b200_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    rc = lib.dalek_b200_init(0, C.byref(h))
    assert rc == -2 and not h.value                     # DALEK_E_NO_DEVICE, no context
    import curve25519_dalek_b200 as pkg
    with pytest.raises(Exception):
        pkg.Engine(0)
    with pytest.raises(Exception):
        pkg.verify_batch([b"m"], [bytes(64)], [bytes(32)])


def test_product_does_not_touch_oracle():
    """No file of the product package or its CUDA sources references oracle/."""
    pkg = os.path.join(ROOT, "curve25519_dalek_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "oracle/" not in txt, f
    r = subprocess.run(["nm", "-D", os.path.join(pkg, "libdalek_b200.so")], capture_output=True, text=True)
    assert "fe_pow2k" not in r.stdout and "msm_pippenger" not in r.stdout


def test_sass_is_sm100a_with_imad_wide():
    r = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "curve25519_dalek_b200", "libdalek_b200.so")],
                       capture_output=True, text=True)
    assert "sm_100a" in r.stdout
