"""The C-ABI library loads without a GPU and exports every symbol include/sdx.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from superdendrix_b200 import _capi, build as sdx_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return sdx_build.build()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "sdx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdx_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_expects():
    assert declared_functions() == sorted(_capi.PROTOTYPES)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), "libsdx.so does not export %s" % name


def test_every_entry_point_cites_the_reference():
    text = open(os.path.join(ROOT, "include", "sdx.h")).read()
    for name in ("sdx_curveball", "sdx_curveball_replay", "sdx_score_sets", "sdx_null_pvalue", "sdx_scan", "sdx_condperm",
                 "sdx_ranksum", "sdx_feature_scores", "sdx_upload_matrix", "sdx_upload_weights"):
        head = text[:text.index("int %s(" % name)]
        comment = head[head.rindex("/*"):]
        assert re.search(r"(src|utils)/[a-z_]+\.py:\d+", comment), "%s has no reference citation" % name


def test_version_and_loud_failure_without_a_gpu(lib_path):
    lib = _capi.load()
    assert lib.sdx_version() >= 100
    import torch
    if not torch.cuda.is_available():
        from superdendrix_b200.engine import Engine, SdxError
        with pytest.raises(SdxError) as e:
            Engine(0)
        assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "superdendrix_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f
