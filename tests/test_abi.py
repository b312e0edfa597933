"""CPU-side checks of the drop-in boundary: the shared library builds, loads and exports every
symbol include/dang_gpu.h declares; without a GPU the product path fails loudly."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dang_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dang_gpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from dang_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = C.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"libdang_gpu.so does not export {s}"
    # the ctypes signature table covers exactly the header
    assert sorted(_lib.SIGNATURES) == syms


def test_library_is_sm100a_and_has_no_oracle_dependency():
    from dang_b200 import _lib
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_product_sources_never_reference_the_oracle():
    pkg = os.path.join(ROOT, "dang_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "dang_oracle" not in src, f


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_fails_loudly_without_a_gpu():
    from dang_b200.engine import DangGpuError, Engine
    from helpers import small_case
    cfg, sky = small_case(nside=4)
    with pytest.raises(DangGpuError) as ei:
        Engine(cfg, sky)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
