"""CPU: the C-ABI shared library loads and exports every symbol include/suta_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "suta_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(suta_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("suta_engine_create", "suta_batch_begin", "suta_forward", "suta_loss_backward", "suta_optimizer_step",
              "suta_reset", "suta_decode", "suta_adapt_step", "suta_op_gemm", "suta_op_loss"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from suta_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/suta_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(declared_symbols())
    assert _lib.load().suta_abi_version() == _lib.ABI_VERSION


def test_ctypes_structs_match_header_sizes():
    from suta_b200 import _lib
    assert ctypes.sizeof(_lib.ModelCfg) == 4 * (6 + 3 * 8 + 2) + 4 + 2 * 4
    assert ctypes.sizeof(_lib.LayerWeights) == 12 * 8
    assert ctypes.sizeof(_lib.Weights) == 8 * (3 + 8 + 8 + 8 + 3 + 3 + 3 + 2) + 48 * 12 * 8
    assert ctypes.sizeof(_lib.Hyper) == 48
    assert ctypes.sizeof(_lib.ParamSeg) == 32


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from suta_b200 import ModelConfig, SutaEngine, _lib
    with pytest.raises(_lib.SutaError):
        SutaEngine(ModelConfig.tiny(), {})
