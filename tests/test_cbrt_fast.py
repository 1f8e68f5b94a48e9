"""CPU check of the ALGORITHM behind the kernels' cube root (common.cuh: msun_cbrtf_fast) against the oracle's cbrtf.

tests/host/cbrt_fast_mirror.c restates the device function in plain C; the two hardware approximations it uses are perturbed by up to
two ulps either way.  The device code itself is compared with the restated msun function on the GPU, on every float of [2^-9, 2)
(tests/test_gpu_parity.py::test_fast_cbrt_is_the_exact_one)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mirror():
    out = os.path.join(tempfile.mkdtemp(prefix="cbrt_mirror_"), "libcbrt_mirror.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host", "cbrt_fast_mirror.c"), "-lm"])
    lib = ctypes.CDLL(out)
    lib.cbrt_compare.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p,
                                 ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long)]
    return lib


def _bits(v):
    return int(np.float32(v).view(np.uint32))


@pytest.mark.parametrize("pert", [0, 2, -2])
def test_fast_cbrt_algorithm_matches_the_oracle(mirror, pert):
    exact = ctypes.cast(ob.lib().ora_cbrtf, ctypes.c_void_p)
    bad, fb, n = ctypes.c_long(0), ctypes.c_long(0), ctypes.c_long(0)
    # every 7th float of [2^-9, 2): 12 M inputs per setting (the opsin transfer feeds it values in [0.0037, 1.004])
    mirror.cbrt_compare(_bits(2.0 ** -9), _bits(2.0), 7, pert, exact, ctypes.byref(bad), ctypes.byref(fb), ctypes.byref(n))
    assert n.value > 11_000_000 and bad.value == 0
    assert 0 < fb.value < n.value * 2e-4        # the exact function is called for ~6 inputs in 100,000
    # the rest of the line: small and large normal numbers on the fast path (below 2^120: the reciprocal of 3x must stay a normal
    # float), and zero, subnormals, the largest floats and negative numbers through the exact function
    for lo, hi in ((0.0, 2.0 ** -9), (2.0, 2.0 ** 119), (2.0 ** 119, np.finfo(np.float32).max), (-1.0, -2.0)):
        mirror.cbrt_compare(_bits(lo), _bits(hi), 1009, pert, exact, ctypes.byref(bad), ctypes.byref(fb), ctypes.byref(n))
        assert n.value > 1000 and bad.value == 0
