"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "whisper_b200.h")).read()
    return sorted(set(re.findall(r"WB_API\s+[\w\s\*]+?\b(wb_\w+)\s*\(", src)))


def test_header_declares_symbols():
    syms = declared_symbols()
    assert len(syms) >= 30
    for must in ("wb_encode", "wb_decode_step", "wb_decode_run", "wb_linear", "wb_layernorm", "wb_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/whisper_b200.h but not exported"


def test_python_signatures_cover_header(lib):
    from whisper_trtllm_b200 import _abi
    assert sorted(_abi.SIGNATURES) == declared_symbols()


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "whisper_b200.h")).read()
    assert "torch" not in src.lower().replace("pytorch", "") or "at::" not in src
    assert "at::Tensor" not in src and "std::" not in src


def test_error_reporting_without_gpu(lib):
    # argument validation happens before any CUDA call: a null config must fail loudly with a message
    out = ctypes.c_void_p()
    rc = lib.wb_model_create(None, 0, ctypes.byref(out))
    assert rc < 0
    assert b"null" in lib.wb_last_error()
    assert lib.wb_version() >= 100


def test_missing_library_fails_loudly(monkeypatch):
    from whisper_trtllm_b200 import _abi
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_abi, "LIB_PATH", "/nonexistent/libwhisper_b200.so")
    with pytest.raises(_abi.WhisperB200Error):
        _abi.load()
