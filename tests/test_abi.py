"""The C-ABI library loads and exports every symbol include/b200rt.h declares (no GPU needed,
no compute calls), and the Python stub types exactly that set."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200rt.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rt_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(binding):
    lib = ctypes.CDLL(binding.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200rt.h but not exported"


def test_python_stub_covers_header(binding):
    assert sorted(binding.SIGNATURES) == declared_symbols()


def test_header_cites_reference():
    src = open(HEADER).read()
    # every entry point documents the reference interface it replaces
    for token in ("RT_gpu.cu", "RT_grid.hpp", "singlet_CFR.hpp", "observation.hpp", "emission_voxels.hpp",
                  "grid_spherical_azimuthally_symmetric.hpp", "boundaries.hpp"):
        assert token in src


def test_no_cpu_fallback(binding):
    """without a CUDA device the product must fail loudly, never compute on the CPU"""
    lib = binding.load()
    if lib.b200rt_device_count() > 0:
        import pytest
        pytest.skip("a GPU is visible here")
    try:
        binding.Context(0)
    except binding.B200RTError as e:
        assert "no usable CUDA device" in str(e)
    else:
        raise AssertionError("Context() succeeded without a GPU")


def test_product_does_not_touch_oracle():
    """the shipped package never imports or links anything under oracle/"""
    pkg = os.path.join(ROOT, "3d_planetary_rt_model_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oraclebind" not in txt and "refbind" not in txt and "rt_oracle" not in txt, f
                assert "/root/reference" not in txt, f
