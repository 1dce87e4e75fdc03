"""ctypes binding of libwhisper_b200.so (C-ABI declared in include/whisper_b200.h).

Loaded the way the reference loads its native plugin library
(tensorrt_llm/plugin/plugin.py:10-22: ``ctypes.CDLL(..., mode=RTLD_GLOBAL)`` + hand-set argtypes/restype).
There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_longlong, c_size_t, c_uint8, c_void_p

LIB_NAME = "libwhisper_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

F32, BF16 = 0, 1


class WhisperB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libwhisper_b200 error {code}: {msg}")
        self.code = code


class wb_config(Structure):
    _fields_ = [(n, c_int32) for n in (
        "d_model", "n_heads", "encoder_layers", "decoder_layers", "ffn_dim", "vocab_size",
        "num_mel_bins", "n_frames", "max_source_positions", "max_target_positions",
        "decoder_start_token_id", "eos_token_id", "pad_token_id", "max_length")]


# name -> (restype, argtypes); every symbol include/whisper_b200.h declares
SIGNATURES = {
    "wb_last_error": (c_char_p, []),
    "wb_version": (c_int, []),
    "wb_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "wb_set_backend": (c_int, [c_int, c_int]),
    "wb_set_pdl": (c_int, [c_int]),
    "wb_set_cuda_graphs": (c_int, [c_int]),
    "wb_set_gemm_block_n": (c_int, [c_int]),
    "wb_set_lean_decode_gemm": (c_int, [c_int]),
    "wb_set_small_batch_path": (c_int, [c_int]),
    "wb_set_decode_chain_path": (c_int, [c_int]),
    "wb_set_step_trace": (c_int, [c_void_p]),
    "wb_set_self_attention_warp_kernel": (c_int, [c_int]),
    "wb_set_decode_attention_backend": (c_int, [c_int]),
    "wb_bandwidth_probe": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p]),
    "wb_launch_count": (c_longlong, []),
    "wb_model_create": (c_int, [POINTER(wb_config), c_int, POINTER(c_void_p)]),
    "wb_model_destroy": (c_int, [c_void_p]),
    "wb_model_load_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "wb_model_set_generation": (c_int, [c_void_p, POINTER(c_int32), c_int, POINTER(c_int32), c_int, c_int, POINTER(c_int32), c_int]),
    "wb_model_weight_bytes": (c_int, [c_void_p, POINTER(c_size_t)]),
    "wb_session_workspace_bytes": (c_int, [c_void_p, c_int, c_int, POINTER(c_size_t)]),
    "wb_session_create": (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t, POINTER(c_void_p)]),
    "wb_session_destroy": (c_int, [c_void_p]),
    "wb_session_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "wb_encode": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "wb_set_encoder_output": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "wb_decode_begin": (c_int, [c_void_p, c_int, c_void_p]),
    "wb_decode_step": (c_int, [c_void_p, c_void_p]),
    "wb_decode_run": (c_int, [c_void_p, c_int, c_int, POINTER(c_int), c_void_p]),
    "wb_decode_run_multi": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, POINTER(c_int), c_void_p]),
    "wb_decode_compact": (c_int, [c_void_p, POINTER(c_int), c_void_p]),
    "wb_decode_refill": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                 c_void_p]),
    "wb_decode_tokens": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int)]),
    "wb_decode_logits": (c_int, [c_void_p, POINTER(c_void_p)]),
    "wb_decode_set_forced_tokens": (c_int, [c_void_p, c_void_p]),
    "wb_decode_set_logits_dump": (c_int, [c_void_p, c_void_p, c_int]),
    "wb_session_cross_kv": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_int64)]),
    "wb_session_self_kv": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), POINTER(c_int)]),
    "wb_session_profile": (c_int, [c_void_p, c_int]),
    "wb_session_profile_at": (c_int, [c_void_p, c_int, c_int]),
    "wb_session_profile_read": (c_int, [c_void_p, POINTER(ctypes.c_double), POINTER(c_longlong)]),
    "wb_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "wb_linear": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                          c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "wb_linear_splitk": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int,
                                 POINTER(c_int), c_void_p]),
    "wb_layernorm_preadd": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_float, c_void_p]),
    "wb_encoder_stem": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "wb_encoder_attention": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "wb_decode_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int64, c_int64, c_void_p]),
    "wb_paged_self_attention": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p]),
    "wb_argmax": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "wb_embed": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "wb_kv_append": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p]),
    "wb_linear_split_heads": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_int, c_void_p]),
    "wb_conv_stem_workspace_bytes": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "wb_conv_stem": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                             c_void_p, c_size_t, c_void_p, c_void_p]),
    "wb_log_mel_workspace_bytes": (c_int, [c_int, POINTER(c_size_t)]),
    "wb_log_mel": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "wb_cast": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p]),
}

_lib = None


def load():
    """Load the shared library once (RTLD_GLOBAL, like plugin.py:13) and set every signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WhisperB200Error(-100, f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                     "or `make -C whisper_trtllm_b200/csrc`; there is no fallback path")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise WhisperB200Error(status, load().wb_last_error().decode("utf-8", "replace"))


def call(name: str, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    check(getattr(load(), name)(*args))


def ptr(t):
    """Raw address of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_handle(stream=None):
    """cudaStream_t as an integer handle — run.py:35-46 passes ``torch.cuda.current_stream().cuda_stream``."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    if isinstance(stream, int):
        return c_void_p(stream)
    return c_void_p(stream.cuda_stream)


__all__ = ["load", "check", "call", "ptr", "stream_handle", "wb_config", "WhisperB200Error", "SIGNATURES", "LIB_PATH",
           "F32", "BF16", "byref", "c_int", "c_void_p", "c_size_t", "c_int32", "c_int64", "c_uint8"]
