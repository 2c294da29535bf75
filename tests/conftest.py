import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from tests.oracle_util import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def hostcheck():
    """TEST-ONLY host compilation of the kernel math (tests/hostcheck/hostcheck.cpp); never loaded by the package."""
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "libhostcheck.so")
    srcs = [os.path.join(d, "hostcheck.cpp")] + [os.path.join(ROOT, "panda_lang_manip_b200", "csrc", f) for f in
                                                 ("panda_dyn.cuh", "panda_contact.cuh", "panda_env.cuh", "panda_model.h", "panda_scene.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, srcs[0]], check=True)
    import ctypes
    lib = ctypes.CDLL(so)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.hc_env_step.argtypes = [ci, ci, ci, ci] + [vp] * 8
    lib.hc_minv.argtypes = [ci] + [vp] * 5
    lib.hc_substeps.argtypes = [ci, vp, vp, vp, vp, ci]
    lib.hc_observe.argtypes = [ci] + [vp] * 6
    lib.hc_ik.argtypes = [ci] + [vp] * 5
    lib.hc_link_state.argtypes = [ci, vp, ci, vp, vp, vp, vp]
    lib.hc_ik_link.argtypes = [ci, vp, ci, vp, vp, vp, vp]
    lib.hc_bare_steps.argtypes = [ci, ci, vp, vp, vp, vp, vp, vp, ci]
    return lib
