"""CPU checks of the C-ABI boundary: the in-tree library builds for sm_100a, loads, exports every
symbol `include/yawb.h` declares, and refuses to run without a GPU (no CPU fallback)."""

import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "yawb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yawb_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from yet_another_wizz_b200 import _lib
    from yet_another_wizz_b200.csrc import build

    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from yet_another_wizz_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/yawb.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert lib.yawb_version() >= 100


def test_built_for_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "yet_another_wizz_b200", "csrc", "libyawb.so")],
                         capture_output=True, text=True)
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback(lib):
    import ctypes

    import yet_another_wizz_b200 as yb

    h = ctypes.c_void_p()
    rc = lib.yawb_create(0, ctypes.byref(h))
    if rc == 0:  # a GPU is present (GPU box): nothing to check here
        lib.yawb_destroy(h)
        pytest.skip("CUDA device present")
    assert b"no CPU fallback" in lib.yawb_last_error()
    with pytest.raises(yb.YawbError):
        yb.Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "yet_another_wizz_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "cpu_port" not in src and "refshim" not in src, f
