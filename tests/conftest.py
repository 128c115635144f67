import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: test needs /root/reference (build container only)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle_clib():
    """compile the oracle's C restatement once (gcc only; no GPU needed)"""
    import subprocess

    so = os.path.join(ORACLE_DIR, "_build", "libpaircount_ref.so")
    src = os.path.join(ORACLE_DIR, "paircount_ref.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    yield
