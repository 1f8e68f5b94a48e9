"""pytest configuration.

`-m "not gpu"`: oracle known-answer tests, host logic, C-ABI symbol checks (no GPU needed).
`-m gpu`: parity tests proper; they call the CUDA path through the C ABI (libsnesgpu.so) and compare it
with the CPU oracle (oracle/), which is test infrastructure only.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def ctx():
    """One libsnesgpu context on cuda:0.  Fails loudly (no skip, no fallback) when the library or the GPU is missing."""
    from snesimage_b200 import engine
    c = engine.Context(0)
    yield c
    c.close()
