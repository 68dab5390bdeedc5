"""CPU checks of the boundary: the C-ABI library loads, exports every symbol the
header declares, and the product path has no CPU fallback and never touches oracle/."""
import ctypes
import os
import re

import pytest
import torch

from gnn_pressure_estimation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gatres_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gatres_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gatres_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared        # the ctypes table binds exactly the header


def test_abi_version_and_param_count():
    lib = _lib.load()
    assert lib.gatres_abi_version() == _lib.ABI_VERSION
    assert lib.gatres_param_count(15, 32) == 65857          # gatres_small
    assert lib.gatres_param_count(25, 128) == 1667585       # gatres_large


def test_model_desc_layout_matches_the_library():
    lib = _lib.load()
    assert lib.gatres_model_desc_bytes() == ctypes.sizeof(_lib.ModelDesc)
    names = [f[0] for f in _lib.ModelDesc._fields_]
    header = open(os.path.join(ROOT, "include", "gatres_b200.h")).read()
    body = header[header.index("typedef struct gatres_model_desc {"):header.index("} gatres_model_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    declared = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[\d+\])?\s*;", body)
    assert declared == names                                   # same fields, same order


def test_argument_validation_reports_errors():
    lib = _lib.load()
    rc = lib.gatres_gat_agg_fwd(None, None, None, None, None, None, None, None, None, 1, 4, 8, 3, 32, 0, None)
    assert rc == -1 and b"unsupported" in lib.gatres_last_error()
    d = _lib.ModelDesc(15, 48, 388, 1, 1246, 0, 8, None, None, None, None, None)
    assert lib.gatres_saved_floats(ctypes.byref(d)) == -1


def test_workspace_sizes():
    lib = _lib.load()
    one = ctypes.c_void_p(16)
    d = _lib.ModelDesc(15, 32, 388, 97, 1246, 0, 32, one, one, one, one, None)
    M = 32 * 388
    layer = M * 32 + 15 * (6 * M * 32 + 12 * M)                 # compact tensors of the layer kernels (SavedLayout)
    # CTA images of the snapshot-resident tensor-core pair (csrc/resident2.cu ImageLayout; 8 CTAs x 49 rows per snapshot):
    # x [R][32], h1 [R][68], y1 2 x [R][32], 4 x a4(2R) scores, h2 [R][36], 4 x a4(R) scores; one x image more at the end
    R, a4 = 49, lambda n: (n + 3) & ~3
    rec = R * 32 + R * 68 + 2 * R * 32 + 4 * a4(2 * R) + R * 36 + 4 * a4(R)
    images = 15 * 32 * 8 * rec + 32 * 8 * R * 32
    assert lib.gatres_saved_floats(ctypes.byref(d)) == max(layer, images) == images
    d128 = _lib.ModelDesc(15, 128, 388, 97, 1246, 0, 32, one, one, one, one, None)
    assert lib.gatres_saved_floats(ctypes.byref(d128)) == M * 128 + 15 * (6 * M * 128 + 12 * M)     # no resident pair at nc = 128
    assert lib.gatres_scratch_floats(ctypes.byref(d), 0) == 8 * M * 32 + 4 * M
    assert lib.gatres_scratch_floats(ctypes.byref(d), 1) == 8 * M * 32 + 10 * M


def test_no_cpu_path():
    import gnn_pressure_estimation_b200.GraphModels as G
    m = G.GATResMeanConv(num_blocks=1, nc=32)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m(torch.zeros(4, 1), torch.zeros(2, 2, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        torch.ops.gatres.apply_mask(torch.zeros(4), torch.zeros(4, dtype=torch.bool))
    with pytest.raises(_lib.GatresError):
        _lib.ptr(torch.zeros(3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gnn_pressure_estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+(torch_geometric|torch_scatter|triton)\b", src, flags=re.M), f
