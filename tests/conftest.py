import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_unavailable_reason():
    """None if the engine can run here (library built, an sm_100 device present), else why not."""
    try:
        import curve25519_dalek_b200 as pkg
        e = pkg.Engine(0)
        e.close()
        return None
    except Exception as exc:                      # missing .so, no CUDA device, not an sm_100 part
        return str(exc)


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a B200 skips the gpu-marked tests instead of erroring in Engine()."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    why = _gpu_unavailable_reason()
    if why is None:
        return
    skip = pytest.mark.skip(reason="needs the B200 engine: " + why)
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        return json.load(f)
