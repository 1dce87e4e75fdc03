"""Layer classes with the names, constructor arguments and parameter attributes of the reference's
``tensorrt_llm.layers`` that the Whisper path touches (layers/linear.py:38-139, layers/normalization.py:6-30,
layers/conv.py, layers/embedding.py, layers/attention.py:216-348), as EAGER modules over CUDA tensors.

In the reference these classes only *describe* a TensorRT network; here ``forward`` executes immediately by calling
the stateless operator entry points of ``libwhisper_b200.so`` (include/whisper_b200.h).  PyTorch only owns the
device memory.  There is no torch arithmetic and no fallback: a forward on a machine without the CUDA library
raises ``WhisperB200Error``.

Conventions (batch generalised from the reference's 1 to B):
  * the residual stream between blocks is fp32; LayerNorm emits the module's compute dtype (fp32 or bf16);
    Linear consumes the compute dtype and emits the compute dtype, or fp32 when a residual is fused in;
  * parameters are kept as fp32 masters (what the binders assign through ``.value``) and packed lazily into the
    compute dtype; the packed copy is refreshed whenever ``.value`` is assigned again.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _abi
from ._abi import c_size_t, byref, ptr, stream_handle

_DT = {torch.float32: _abi.F32, torch.bfloat16: _abi.BF16}


def _torch_dtype(dtype) -> torch.dtype:
    if dtype is None:
        return torch.float32
    if isinstance(dtype, torch.dtype):
        if dtype not in _DT:
            raise ValueError(f"unsupported dtype {dtype}: the path computes in float32 or bfloat16")
        return dtype
    s = str(dtype).lower()
    if s in ("float32", "fp32", "f32"):
        return torch.float32
    if s in ("bfloat16", "bf16"):
        return torch.bfloat16
    raise ValueError(f"unsupported dtype {dtype}: the path computes in float32 or bfloat16")


def default_device() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _abi.WhisperB200Error(-101, f"{what} must be a CUDA tensor: libwhisper_b200 has no CPU path")


class Parameter(torch.nn.Parameter):
    """``tensorrt_llm.parameter.Parameter``: binders assign numpy arrays through ``.value``
    (build_decoder.py:72-101).  Stored as an fp32 master on the module's device."""

    def __new__(cls, shape=None, dtype=None, value=None, device=None):
        if value is None:
            value = torch.zeros(tuple(shape), dtype=torch.float32, device=device or default_device())
        p = super().__new__(cls, value, requires_grad=False)
        p._wb_version = 0
        return p

    @property
    def value(self) -> torch.Tensor:
        return self.data

    @value.setter
    def value(self, v):
        t = torch.from_numpy(np.array(v, dtype=np.float32)) if isinstance(v, np.ndarray) else torch.as_tensor(v)
        self.data = t.detach().to(device=self.data.device, dtype=torch.float32).contiguous()
        self._wb_version = getattr(self, "_wb_version", 0) + 1


class Module(torch.nn.Module):
    """``tensorrt_llm.module.Module`` stand-in: a torch module with a compute dtype and a cache of packed weights."""

    def __init__(self):
        super().__init__()
        self._packed = {}

    def _pack(self, key, params, fn):
        """Cache ``fn()`` until one of ``params`` is re-assigned (or moved)."""
        sig = tuple((p.data_ptr(), getattr(p, "_wb_version", 0)) for p in params if p is not None)
        hit = self._packed.get(key)
        if hit is None or hit[0] != sig:
            hit = (sig, fn())
            self._packed[key] = hit
        return hit[1]


class ModuleList(torch.nn.ModuleList):
    pass


def _as_compute(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Cast an activation to the compute dtype with the library's cast kernel (no torch arithmetic)."""
    require_cuda(x, "activation")
    x = x if x.is_contiguous() else x.contiguous()
    if x.dtype == dtype:
        return x
    if x.dtype not in _DT:
        raise ValueError(f"activation dtype {x.dtype} not supported")
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    _abi.call("wb_cast", ptr(x), _DT[x.dtype], ptr(out), _DT[dtype], x.numel(), stream_handle())
    return out


class LayerNorm(Module):
    """layers/normalization.py:6-30 (eps 1e-5 as the oracle's nn.LayerNorm)."""

    def __init__(self, normalized_shape, eps=1e-05, elementwise_affine=True, dtype=None):
        super().__init__()
        d = normalized_shape if isinstance(normalized_shape, int) else normalized_shape[-1]
        self.normalized_shape = d
        self.eps = eps
        self.dtype = _torch_dtype(dtype)
        self.weight = Parameter(value=torch.ones(d, device=default_device()))
        self.bias = Parameter(value=torch.zeros(d, device=default_device()))

    def forward(self, x: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        x = _as_compute(x, torch.float32)
        d = self.normalized_shape
        assert x.shape[-1] == d
        out = torch.empty(x.shape, dtype=out_dtype or self.dtype, device=x.device)
        _abi.call("wb_layernorm", ptr(x), ptr(self.weight.data), ptr(self.bias.data), ptr(out), _DT[out.dtype],
                  x.numel() // d, d, self.eps, stream_handle())
        return out


class ColumnLinear(Module):
    """layers/linear.py:38-95: y = x W^T + b with W [out_features, in_features] (tp_size 1)."""

    def __init__(self, in_features, out_features, bias=True, dtype=None, tp_group=None, tp_size=1, gather_output=True):
        super().__init__()
        if tp_size != 1:
            raise ValueError("tensor parallelism is not part of this path (utterances are sharded data-parallel)")
        self.in_features, self.out_features = in_features, out_features
        self.dtype = _torch_dtype(dtype)
        self.weight = Parameter(shape=(out_features, in_features))
        if bias:
            self.bias = Parameter(shape=(out_features,))
        else:
            self.register_parameter("bias", None)

    def packed_weight(self, scale: float = 1.0) -> torch.Tensor:
        return self._pack(("w", self.dtype, scale), [self.weight],
                          lambda: (self.weight.data * scale if scale != 1.0 else self.weight.data).to(self.dtype).contiguous())

    def packed_bias(self, scale: float = 1.0) -> Optional[torch.Tensor]:
        if self.bias is None:
            return None
        return self._pack(("b", scale), [self.bias], lambda: (self.bias.data * scale).contiguous())

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None, act: int = 0,
                out_dtype: Optional[torch.dtype] = None, scale: float = 1.0) -> torch.Tensor:
        """``scale`` (a power of two) is folded into the packed weight and bias — used for the q projection
        (q * head_dim**-0.5, modeling_whisper.py:472; exact because 0.125 commutes with every rounding)."""
        x = _as_compute(x, self.dtype)
        K, N = self.in_features, self.out_features
        assert x.shape[-1] == K, (x.shape, K)
        M = x.numel() // K
        if residual is not None:
            residual = _as_compute(residual, torch.float32)
            out_dtype = torch.float32
        out = torch.empty(*x.shape[:-1], N, dtype=out_dtype or self.dtype, device=x.device)
        _abi.call("wb_linear", ptr(x), K, ptr(self.packed_weight(scale)), K, _DT[self.dtype], ptr(self.packed_bias(scale)),
                  ptr(residual), N if residual is not None else 0, ptr(out), N, _DT[out.dtype], M, N, K, act, 0, stream_handle())
        return out


class RowLinear(ColumnLinear):
    """layers/linear.py:98-139 (identical to ColumnLinear at tp_size 1)."""

    def __init__(self, in_features, out_features, bias=True, dtype=None, tp_group=None, tp_size=1):
        super().__init__(in_features, out_features, bias=bias, dtype=dtype, tp_group=tp_group, tp_size=tp_size)


Linear = ColumnLinear


class Embedding(Module):
    """layers/embedding.py: a [num_embeddings, embedding_dim] table (lookups are fused into wb_embed)."""

    def __init__(self, num_embeddings, embedding_dim, dtype=None):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.dtype = _torch_dtype(dtype)
        self.weight = Parameter(shape=(num_embeddings, embedding_dim))

    def packed_weight(self) -> torch.Tensor:
        return self._pack(("w", self.dtype), [self.weight], lambda: self.weight.data.to(self.dtype).contiguous())


class Conv2d(Module):
    """layers/conv.py as the reference uses it: kernel (1,3) because TRT-LLM lacked Conv1d (model.py:77-79).
    The weight is [out, in, 1, 3]; execution happens in the fused stem (``WhisperEncoder.forward`` -> wb_conv_stem)."""

    def __init__(self, in_channels, out_channels, kernel_size=(1, 3), stride=(1, 1), padding=(0, 1), dtype=None):
        super().__init__()
        if tuple(kernel_size) != (1, 3) or tuple(padding) != (0, 1) or tuple(stride) not in ((1, 1), (1, 2)):
            raise ValueError("the Whisper stem uses kernel (1,3), padding (0,1), stride (1,1) or (1,2)")
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, tuple(stride)
        self.dtype = _torch_dtype(dtype)
        self.weight = Parameter(shape=(out_channels, in_channels, 1, 3))
        self.bias = Parameter(shape=(out_channels,))

    def packed_weight(self, kpad: Optional[int] = None) -> torch.Tensor:
        """[out, in, 1, 3] -> [out, 3*in (zero padded to kpad)] with k = tap * in + c (the GEMM-stem layout)."""
        def build():
            w = self.weight.data.reshape(self.out_channels, self.in_channels, 3).permute(0, 2, 1).reshape(self.out_channels, -1)
            if kpad is not None and kpad > w.shape[1]:
                w = torch.nn.functional.pad(w, (0, kpad - w.shape[1]))
            return w.to(self.dtype).contiguous()
        return self._pack(("w", self.dtype, kpad), [self.weight], build)


class Attention(Module):
    """The stock ``Attention`` layer in its no-plugin branch (layers/attention.py:216-348) as the encoder uses it:
    fused ``qkv`` ColumnLinear (zero K bias, build_encoder.py:78-79) + ``dense``; bidirectional, no mask,
    softmax in fp32.  scores = (q k^T) / sqrt(head_dim): the 1/8 is folded into the packed q rows (exact)."""

    def __init__(self, hidden_size, num_attention_heads, num_layers=1, dtype=None, **_unused):
        super().__init__()
        self.hidden_size, self.num_attention_heads = hidden_size, num_attention_heads
        self.attention_head_size = hidden_size // num_attention_heads
        if self.attention_head_size != 64:
            raise ValueError("head_dim must be 64 (all Whisper sizes)")
        self.dtype = _torch_dtype(dtype)
        self.qkv = ColumnLinear(hidden_size, 3 * hidden_size, bias=True, dtype=dtype)
        self.dense = RowLinear(hidden_size, hidden_size, bias=True, dtype=dtype)

    def _packed_qkv(self):
        d = self.hidden_size
        def build():
            w = self.qkv.weight.data.clone()
            b = self.qkv.bias.data.clone()
            w[:d] *= 0.125
            b[:d] *= 0.125
            return w.to(self.dtype).contiguous(), b.contiguous()
        return self._pack(("qkv", self.dtype), [self.qkv.weight, self.qkv.bias], build)

    def forward(self, hidden_states: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = _as_compute(hidden_states, self.dtype)
        B, S, d = x.shape
        w, b = self._packed_qkv()
        qkv = torch.empty(B * S, 3 * d, dtype=self.dtype, device=x.device)
        _abi.call("wb_linear", ptr(x), d, ptr(w), d, _DT[self.dtype], ptr(b), None, 0, ptr(qkv), 3 * d, _DT[self.dtype],
                  B * S, 3 * d, d, 0, 0, stream_handle())
        ctx = torch.empty(B, S, d, dtype=self.dtype, device=x.device)
        _abi.call("wb_encoder_attention", ptr(qkv), ptr(ctx), _DT[self.dtype], B, S, self.num_attention_heads, 0, stream_handle())
        return self.dense(ctx, residual=residual)


def conv_stem(conv1: Conv2d, conv2: Conv2d, positions: torch.Tensor, mel: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """gelu(conv1) -> gelu(conv2, stride 2) -> permute -> + positions   (model.py:94-102), fp32 [B, n_frames/2, d]."""
    require_cuda(mel, "input_features")
    mel = _as_compute(mel, torch.float32)
    B, n_mels, T = mel.shape
    d = conv1.out_channels
    nbytes = c_size_t()
    _abi.call("wb_conv_stem_workspace_bytes", B, d, T, _DT[dtype], byref(nbytes))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=mel.device)
    x = torch.empty(B, T // 2, d, dtype=torch.float32, device=mel.device)
    pos = positions.reshape(T // 2, d).contiguous()
    _abi.call("wb_conv_stem", ptr(mel), B, ptr(conv1.packed_weight(256)), ptr(conv1.bias.data), ptr(conv2.packed_weight()),
              ptr(conv2.bias.data), ptr(pos), _DT[dtype], d, n_mels, T, ptr(ws), c_size_t(nbytes.value), ptr(x), stream_handle())
    return x
