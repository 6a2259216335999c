"""The C-ABI library builds, loads and exports every symbol include/drb.h declares
(no compute calls: there is no GPU in the CPU test run)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib_path():
    from differentiable_ransac_b200 import build

    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "drb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/drb.h but not exported"


def test_binding_table_matches_header():
    from differentiable_ransac_b200 import _lib

    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_status_strings_and_version(lib_path):
    from differentiable_ransac_b200 import _lib

    lib = _lib.load()
    assert lib.drb_version() >= 100
    assert lib.drb_status_string(0) == b"ok"
    assert lib.drb_status_string(-2) == b"bad shape"


def test_null_and_shape_errors_do_not_need_a_gpu(lib_path):
    from differentiable_ransac_b200 import _lib

    lib = _lib.load()
    assert lib.drb_sample(None, None, 0, 0, None, 1.0, 1, 1, 8, 5, None, None, None, None, None) == -1
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.drb_sample(p, None, 0, 0, None, 1.0, 1, 1, 4, 5, p, None, None, None, None) == -2      # s > N
    assert lib.drb_sample(p, None, 0, 0, None, 1.0, 1, 1, 8, 6, p, None, None, None, None) == -3      # s unsupported
    assert lib.drb_solve_e5(None, None, 1, 1, 1, None, None, None, None, None, None) == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "differentiable_ransac_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "hostcheck" not in src, f
