"""Helpers for the -m gpu tests: call the C-ABI operators on torch CUDA tensors."""
import torch

from whisper_trtllm_b200 import _abi
from whisper_trtllm_b200._abi import ptr, stream_handle

DT = {torch.float32: _abi.F32, torch.bfloat16: _abi.BF16}


def linear(A, W, bias=None, residual=None, act=0, out_dtype=torch.float32, backend=0, lda=None, M=None, K=None):
    """out = act(A W^T + bias) + residual via wb_linear."""
    M = A.shape[0] if M is None else M
    K = A.shape[1] if K is None else K
    N = W.shape[0]
    out = torch.empty(M, N, dtype=out_dtype, device=A.device)
    _abi.call("wb_linear", ptr(A), A.stride(0) if lda is None else lda, ptr(W), W.stride(0), DT[A.dtype], ptr(bias),
              ptr(residual), residual.stride(0) if residual is not None else 0, ptr(out), N, DT[out_dtype], M, N, K, act,
              backend, stream_handle())
    return out


def layernorm(x, g, b, out_dtype=torch.float32, eps=1e-5):
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    _abi.call("wb_layernorm", ptr(x), ptr(g), ptr(b), ptr(out), DT[out_dtype], x.shape[0], x.shape[1], eps, stream_handle())
    return out


def encoder_attention(qkv, B, S, H, backend=0):
    out = torch.empty(B * S, H * 64, dtype=qkv.dtype, device=qkv.device)
    _abi.call("wb_encoder_attention", ptr(qkv), ptr(out), DT[qkv.dtype], B, S, H, backend, stream_handle())
    return out


def decode_attention(q, k, v, n_keys):
    B, H, T, _ = k.shape
    out = torch.empty(B, H * 64, dtype=q.dtype, device=q.device)
    _abi.call("wb_decode_attention", ptr(q), ptr(k), ptr(v), ptr(out), DT[q.dtype], B, H, n_keys, k.stride(0), k.stride(1),
              stream_handle())
    return out


def argmax(logits, mask=None, bits=1):
    out = torch.empty(logits.shape[0], dtype=torch.int32, device=logits.device)
    _abi.call("wb_argmax", ptr(logits), logits.stride(0), logits.shape[0], logits.shape[1], ptr(mask), bits, ptr(out),
              stream_handle())
    return out


def rel_err(a, b):
    """max |a-b| relative to max |b| (the 'relative-to-max' measure used for the stated tolerances)."""
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
