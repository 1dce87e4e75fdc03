"""Drop-in model classes: same names, constructor kwargs, parameter attribute names and forward signatures as
``tensorrt_llm/models/whisper/model.py`` of the reference (WhisperEncoder :68-124, WhisperDecoderAttention :153-304,
WhisperDecoderLayer :306-369, WhisperDecoder :371-516), as eager modules over CUDA tensors that run on the
hand-written sm_100a kernels of libwhisper_b200.so.

Differences that are part of the contract:
  * plain tensors instead of RaggedTensor; the batch dimension is generalised from 1 to B;
  * numerics follow the ORACLE where the two reference implementations differ: erf GELU,
    q scaled before q k^T (folded into the packed q weights, exact), no K bias;
  * ``cache_mask`` carries the cache length in its SHAPE (``cache_mask.shape[0] - 1``); its values are never read
    (the reference's runner fills it with ``torch.rand``, run.py:118-126).

The fast transcription path does not go through these per-layer calls: ``WhisperEngine`` / ``run.greedy_search``
pack the same weights into the native runtime (paged in-place KV cache, on-device greedy loop).  These classes are
the module-level drop-ins and the place where the four attention modes are tested one by one.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _abi
from ._abi import ptr, stream_handle
from .layers import (Attention, ColumnLinear, Conv2d, Embedding, LayerNorm, Module, ModuleList, RowLinear, _DT, _as_compute,
                     _torch_dtype, conv_stem, default_device, require_cuda)

ACT2FN = {"gelu": 1}  # epilogue code of wb_linear: exact-erf GELU (transformers/activations.py:214)


def _act_code(activation_function: str) -> int:
    if activation_function not in ACT2FN:
        raise ValueError(f"activation {activation_function!r} is not on the Whisper path (only 'gelu')")
    return ACT2FN[activation_function]


class WhisperEncoderLayer(Module):
    """model.py:36-66 — pre-LN self-attention + residual, pre-LN erf-GELU MLP + residual."""

    def __init__(self, d_model=512, encoder_attention_heads=8, activation_function="gelu", encoder_ffn_dim=2048, dtype=None):
        super().__init__()
        self.embed_dim = d_model
        self.self_attn = Attention(self.embed_dim, encoder_attention_heads, 1, dtype=dtype)
        self.self_attn_layer_norm = LayerNorm(self.embed_dim, dtype=dtype)
        self.activation_fn = _act_code(activation_function)
        self.fc1 = ColumnLinear(self.embed_dim, encoder_ffn_dim, dtype=dtype)
        self.fc2 = ColumnLinear(encoder_ffn_dim, self.embed_dim, dtype=dtype)
        self.final_layer_norm = LayerNorm(self.embed_dim, dtype=dtype)

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        residual = _as_compute(hidden_states, torch.float32)
        h = self.self_attn_layer_norm(residual)
        residual = self.self_attn(h, residual=residual)          # residual + attn, fused in the dense epilogue
        h = self.final_layer_norm(residual)
        h = self.fc1(h, act=self.activation_fn)
        return self.fc2(h, residual=residual)


class WhisperEncoder(Module):
    """model.py:68-124.  ``forward(input_features [B, 80, 3000]) -> hidden_states fp32 [B, 1500, d]``."""

    def __init__(self, d_model=512, num_mel_bins=80, max_source_positions=1500, encoder_layers=6, encoder_attention_heads=8,
                 activation_function="gelu", encoder_ffn_dim=2048, dtype=None):
        super().__init__()
        embed_dim = d_model
        self.d_model, self.num_mel_bins, self.max_source_positions = d_model, num_mel_bins, max_source_positions
        self.dtype = _torch_dtype(dtype)
        # Conv1d in the oracle (modeling_whisper.py:934-935); the reference spells it Conv2d with a (1,3) kernel
        self.conv1 = Conv2d(num_mel_bins, embed_dim, kernel_size=(1, 3), padding=(0, 1), dtype=dtype)
        self.conv2 = Conv2d(embed_dim, embed_dim, kernel_size=(1, 3), stride=(1, 2), padding=(0, 1), dtype=dtype)
        self.embed_positions_weight = torch.zeros(1, max_source_positions, embed_dim).numpy()
        self.layers = ModuleList([WhisperEncoderLayer(d_model=d_model, encoder_attention_heads=encoder_attention_heads,
                                                      activation_function=activation_function,
                                                      encoder_ffn_dim=encoder_ffn_dim, dtype=dtype)
                                  for _ in range(encoder_layers)])
        self.layer_norm = LayerNorm(embed_dim, dtype=dtype)
        self._pos_cache = None

    def _positions(self, device) -> torch.Tensor:
        src = self.embed_positions_weight
        key = (id(src), str(device))
        if self._pos_cache is None or self._pos_cache[0] != key:
            t = torch.from_numpy(np.array(src, copy=True)) if isinstance(src, np.ndarray) else torch.as_tensor(src)   # own, writable copy
            self._pos_cache = (key, t.detach().to(device=device, dtype=torch.float32).contiguous(), src)
        return self._pos_cache[1]

    def forward(self, input_features: torch.Tensor) -> torch.Tensor:
        require_cuda(input_features, "input_features")
        h = conv_stem(self.conv1, self.conv2, self._positions(input_features.device), input_features, self.dtype)
        for layer in self.layers:
            h = layer(h)
        return self.layer_norm(h, out_dtype=torch.float32)   # the engine output 'hidden_states' is fp32 (model.py:109)

    def prepare_inputs(self, batch_size: int = 1):
        """Example inputs with the engine's tensor contract (model.py:113-124): data f32 [B,80,3000] (+ unused length)."""
        dev = default_device()
        return torch.rand(batch_size, self.num_mel_bins, 2 * self.max_source_positions, device=dev)


class WhisperDecoderAttention(Module):
    """model.py:153-304.  ``forward(hidden_states, key_value_states=None, past_key=None, past_value=None,
    cache_mask=None) -> (context, past_key, past_value)`` with the reference's four modes:

      self  / no cache   cache_mask.shape[0]-1 == 0 (or past_key None): K,V = proj(hidden)             (:273-281)
      self  / cache      K = concat(past_key[:, :, :n], proj(hidden)), n = min(len(mask)-1, past_key.shape[2])
      cross / no cache   cache_mask.shape[0]-1 == 0: K,V = proj(key_value_states)                       (:261-272)
      cross / cache      cache_mask.shape[0]-1 == 1500: K,V = past (returned as the same storage, zero copy)

    hidden_states [B, 1, d]; caches [B, H, T, 64]; one query token (the greedy loop never feeds more).
    """

    def __init__(self, hidden_size=512, num_attention_heads=8, max_position_embeddings=0, num_layers=1,
                 apply_query_key_layer_scaling=False, bias=True, dtype=None, tp_group=None, tp_size=1, **_unused):
        super().__init__()
        if tp_size != 1 or apply_query_key_layer_scaling:
            raise ValueError("tensor parallelism / layer scaling are not part of the Whisper path")
        self.attention_head_size = hidden_size // num_attention_heads
        if self.attention_head_size != 64:
            raise ValueError("head_dim must be 64 (all Whisper sizes)")
        self.num_attention_heads = num_attention_heads
        self.hidden_size = hidden_size
        self.norm_factor = math.sqrt(self.attention_head_size)
        self.dtype = _torch_dtype(dtype)
        self.q_proj = ColumnLinear(hidden_size, hidden_size, bias=bias, dtype=dtype)
        self.k_proj = ColumnLinear(hidden_size, hidden_size, bias=False, dtype=dtype)
        self.v_proj = ColumnLinear(hidden_size, hidden_size, bias=bias, dtype=dtype)
        self.dense = RowLinear(hidden_size, hidden_size, bias=bias, dtype=dtype)

    # -- helpers ------------------------------------------------------------------------------------------
    def _split_heads_proj(self, lin: ColumnLinear, x: torch.Tensor) -> torch.Tensor:
        """transpose_for_scores(lin(x)) for x [B, S, d] -> [B, H, S, 64], transposed in the GEMM epilogue."""
        B, S, d = x.shape
        out = torch.empty(B, self.num_attention_heads, S, 64, dtype=self.dtype, device=x.device)
        _abi.call("wb_linear_split_heads", ptr(x), d, ptr(lin.packed_weight()), d, _DT[self.dtype], ptr(lin.packed_bias()),
                  ptr(out), _DT[self.dtype], B, S, self.num_attention_heads, d, stream_handle())
        return out

    def _append(self, past: Optional[torch.Tensor], n: int, cur: torch.Tensor) -> torch.Tensor:
        B, H = cur.shape[0], self.num_attention_heads
        out = torch.empty(B, H, n + 1, 64, dtype=self.dtype, device=cur.device)
        if n > 0:
            past = _as_compute(past, self.dtype) if past.dtype != self.dtype else past
            assert past.stride(3) == 1 and past.stride(2) == 64, "past cache rows must be contiguous [.., T, 64]"
            pb, ph = past.stride(0), past.stride(1)
        else:
            pb = ph = 0
        _abi.call("wb_kv_append", ptr(past) if n > 0 else None, pb, ph, n, ptr(cur), cur.stride(0), ptr(out), _DT[self.dtype],
                  B, H, stream_handle())
        return out

    def _attend(self, q: torch.Tensor, key: torch.Tensor, value: torch.Tensor) -> torch.Tensor:
        B, H, n = key.shape[0], self.num_attention_heads, key.shape[2]
        assert key.stride(3) == 1 and key.stride(2) == 64 and value.stride() == key.stride()
        ctx = torch.empty(B, H * 64, dtype=self.dtype, device=q.device)
        _abi.call("wb_decode_attention", ptr(q), ptr(key), ptr(value), ptr(ctx), _DT[self.dtype], B, H, n, key.stride(0),
                  key.stride(1), stream_handle())
        return ctx

    # -- forward --------------------------------------------------------------------------------------------
    def forward(self, hidden_states: torch.Tensor, key_value_states: Optional[torch.Tensor] = None,
                past_key: Optional[torch.Tensor] = None, past_value: Optional[torch.Tensor] = None,
                cache_mask: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None):
        require_cuda(hidden_states, "hidden_states")
        x = _as_compute(hidden_states, self.dtype)
        if x.dim() == 2:
            x = x.unsqueeze(1)
        B, T, d = x.shape
        if T != 1:
            raise ValueError("WhisperDecoderAttention attends with ONE query token per row (greedy decode); got T=%d" % T)
        mask_len = int(cache_mask.shape[0]) - 1 if cache_mask is not None else (0 if past_key is None else None)
        q = self.q_proj(x, scale=0.125).view(B, d)            # q * head_dim**-0.5, exact (power of two)
        if key_value_states is not None:                      # ---- cross attention
            kvs = _as_compute(key_value_states, self.dtype)
            S = kvs.shape[1]
            cache_length = S if mask_len is None else mask_len
            if cache_length < 0 or cache_length > S:
                raise ValueError(f"cross cache length {cache_length} out of range [0, {S}]")
            if cache_length == S:                             # cross / cache: reuse, zero copy
                key = past_key if past_key.dtype == self.dtype else _as_compute(past_key, self.dtype)
                value = past_value if past_value.dtype == self.dtype else _as_compute(past_value, self.dtype)
                if key.stride(2) != 64 or key.stride(3) != 1:
                    key, value = key.contiguous(), value.contiguous()
            elif cache_length == 0:                           # cross / no cache: project the encoder states
                key = self._split_heads_proj(self.k_proj, kvs)
                value = self._split_heads_proj(self.v_proj, kvs)
            else:                                             # reference's general slice+concat form (:263-272)
                cur = kvs[:, :S - cache_length].contiguous()
                key = torch.cat([_as_compute(past_key, self.dtype)[:, :, :cache_length], self._split_heads_proj(self.k_proj, cur)], dim=2)
                value = torch.cat([_as_compute(past_value, self.dtype)[:, :, :cache_length], self._split_heads_proj(self.v_proj, cur)], dim=2)
        else:                                                 # ---- self attention
            avail = 0 if past_key is None else int(past_key.shape[2])
            cache_length = avail if mask_len is None else min(mask_len, avail)
            x2 = x.view(B, d)
            k_cur = self.k_proj(x2)
            v_cur = self.v_proj(x2)
            key = self._append(past_key, cache_length, k_cur)
            value = self._append(past_value, cache_length, v_cur)
        ctx = self._attend(q, key, value)
        context = self.dense(ctx.view(B, 1, d), residual=residual)
        return context, key, value


class WhisperDecoderLayer(Module):
    """model.py:306-369."""

    def __init__(self, d_model=512, decoder_attention_heads=8, activation_function="gelu", decoder_ffn_dim=2048, dtype=None):
        super().__init__()
        self.embed_dim = d_model
        self.self_attn = WhisperDecoderAttention(self.embed_dim, decoder_attention_heads, dtype=dtype)
        self.activation_fn = _act_code(activation_function)
        self.self_attn_layer_norm = LayerNorm(self.embed_dim, dtype=dtype)
        self.encoder_attn = WhisperDecoderAttention(self.embed_dim, decoder_attention_heads, dtype=dtype)
        self.encoder_attn_layer_norm = LayerNorm(self.embed_dim, dtype=dtype)
        self.fc1 = ColumnLinear(self.embed_dim, decoder_ffn_dim, dtype=dtype)
        self.fc2 = ColumnLinear(decoder_ffn_dim, self.embed_dim, dtype=dtype)
        self.final_layer_norm = LayerNorm(self.embed_dim, dtype=dtype)

    def forward(self, hidden_states, encoder_hidden_states=None, self_past_key=None, self_past_value=None,
                self_cache_mask=None, cross_past_key=None, cross_past_value=None, cross_cache_mask=None):
        residual = _as_compute(hidden_states, torch.float32)
        h = self.self_attn_layer_norm(residual)
        residual, present_key, present_value = self.self_attn(
            hidden_states=h, key_value_states=None, past_key=self_past_key, past_value=self_past_value,
            cache_mask=self_cache_mask, residual=residual)
        h = self.encoder_attn_layer_norm(residual)
        residual, cross_key, cross_value = self.encoder_attn(
            hidden_states=h, key_value_states=encoder_hidden_states, past_key=cross_past_key, past_value=cross_past_value,
            cache_mask=cross_cache_mask, residual=residual)
        h = self.final_layer_norm(residual)
        h = self.fc1(h, act=self.activation_fn)
        h = self.fc2(h, residual=residual)
        return h, present_key, present_value, cross_key, cross_value


class WhisperDecoder(Module):
    """model.py:371-516.  ``forward(input_ids, encoder_hidden_states, past_self_keys, past_self_values,
    past_cross_keys, past_cross_values, past_self_cache_mask, past_cross_cache_mask)
    -> (logits, next_self_keys, next_self_values, next_cross_keys, next_cross_values)``.

    Stacked caches: the reference's ``[L, H, T, 64]`` (batch 1) or, batched, ``[L, B, H, T, 64]``; outputs come back in
    the rank they were given.  ``logits`` is fp32 [B, 1, vocab] (the engine output named 'hidden_states', :464)."""

    def __init__(self, pad_token_id=50256, max_target_positions=448, max_source_positions=1500, d_model=512,
                 scale_embedding=False, vocab_size=51864, decoder_layers=6, decoder_attention_heads=8,
                 activation_function="gelu", decoder_ffn_dim=2048, dtype=None):
        super().__init__()
        self.padding_idx = pad_token_id
        self.max_target_positions = max_target_positions
        self.max_source_positions = max_source_positions
        self.d_model = d_model
        self.vocab_size = vocab_size
        # the oracle's forward never multiplies by embed_scale (modeling_whisper.py:1143-1156); neither does model.py:423
        self.embed_scale = math.sqrt(d_model) if scale_embedding else 1.0
        self.decoder_layers = decoder_layers
        self.decoder_attention_heads = decoder_attention_heads
        self.d_head = d_model // decoder_attention_heads
        self.dtype = _torch_dtype(dtype)
        self.embed_tokens = Embedding(vocab_size, d_model, dtype=dtype)
        self.embed_positions = Embedding(self.max_target_positions, d_model, dtype=dtype)
        self.layers = ModuleList([WhisperDecoderLayer(d_model=d_model, decoder_attention_heads=decoder_attention_heads,
                                                      activation_function=activation_function,
                                                      decoder_ffn_dim=decoder_ffn_dim, dtype=dtype)
                                  for _ in range(decoder_layers)])
        self.layer_norm = LayerNorm(d_model, dtype=dtype)
        self.proj_out = ColumnLinear(d_model, vocab_size, bias=False, dtype=dtype)

    def forward(self, input_ids, encoder_hidden_states, past_self_keys, past_self_values, past_cross_keys, past_cross_values,
                past_self_cache_mask, past_cross_cache_mask):
        require_cuda(encoder_hidden_states, "encoder_hidden_states")
        dev = encoder_hidden_states.device
        ids = input_ids.to(device=dev, dtype=torch.int32).contiguous()
        if ids.dim() == 1:
            ids = ids.unsqueeze(1)
        B, T = ids.shape
        position = int(past_self_cache_mask.shape[0]) - 1     # model.py:424
        x = torch.empty(B, T, self.d_model, dtype=torch.float32, device=dev)
        _abi.call("wb_embed", ptr(ids), ids.stride(0), B, T, position, ptr(self.embed_tokens.packed_weight()),
                  ptr(self.embed_positions.packed_weight()), _DT[self.dtype], self.d_model, self.vocab_size, ptr(x), stream_handle())

        stacked4 = past_self_keys is not None and past_self_keys.dim() == 4   # reference layout [L, H, T, 64] (B == 1)
        def layer_slice(t, idx):
            if t is None:
                return None
            s = t[idx]
            return s.unsqueeze(0) if t.dim() == 4 else s
        nsk, nsv, nck, ncv = [], [], [], []
        for idx, layer in enumerate(self.layers):
            x, sk, sv, ck, cv = layer(
                hidden_states=x, encoder_hidden_states=encoder_hidden_states,
                self_past_key=layer_slice(past_self_keys, idx), self_past_value=layer_slice(past_self_values, idx),
                self_cache_mask=past_self_cache_mask,
                cross_past_key=layer_slice(past_cross_keys, idx), cross_past_value=layer_slice(past_cross_values, idx),
                cross_cache_mask=past_cross_cache_mask)
            nsk.append(sk); nsv.append(sv); nck.append(ck); ncv.append(cv)
        h = self.layer_norm(x)
        logits = self.proj_out(h, out_dtype=torch.float32)
        def stack(ts):
            s = torch.stack(ts, dim=0)                        # [L, B, H, T, 64]
            return s[:, 0] if stacked4 else s
        # cross cache reuse: when every layer handed back the caller's own storage, return the caller's tensors
        def stack_cross(ts, given):
            if given is not None and all(t.data_ptr() == layer_slice(given, i).data_ptr() for i, t in enumerate(ts)):
                return given
            return stack(ts)
        return logits, stack(nsk), stack(nsv), stack_cross(nck, past_cross_keys), stack_cross(ncv, past_cross_values)

    def prepare_inputs(self, batch_size: int = 1, past_length: int = 0):
        """Example inputs with the engine's tensor contract (model.py:472-516, run.py:105-126): step 0 passes
        dummy caches with masks of length 1; afterwards masks of length past+1 / 1500+1."""
        dev = default_device()
        L, H, dh, S = self.decoder_layers, self.decoder_attention_heads, self.d_head, self.max_source_positions
        shape = lambda t: (L, batch_size, H, t, dh) if batch_size > 1 else (L, H, t, dh)
        ids = torch.zeros(batch_size, 1, dtype=torch.int32, device=dev)
        enc = torch.rand(batch_size, S, self.d_model, device=dev)
        if past_length == 0:
            return (ids, enc, torch.rand(shape(1), device=dev), torch.rand(shape(1), device=dev), torch.rand(shape(S), device=dev),
                    torch.rand(shape(S), device=dev), torch.rand(1, device=dev), torch.rand(1, device=dev))
        return (ids, enc, torch.rand(shape(past_length), device=dev), torch.rand(shape(past_length), device=dev),
                torch.rand(shape(S), device=dev), torch.rand(shape(S), device=dev), torch.rand(past_length + 1, device=dev),
                torch.rand(S + 1, device=dev))


# ------------------------------------------------------------------------------------------------------------
# binders: HF state_dict -> module parameters, the mapping of build_encoder.py:71-91 / build_decoder.py:71-101
# ------------------------------------------------------------------------------------------------------------
def _np(t):
    return t.detach().to("cpu", torch.float32).numpy() if isinstance(t, torch.Tensor) else np.asarray(t, dtype=np.float32)


def load_encoder_from_hf(model: WhisperEncoder, ckpt: Dict[str, torch.Tensor]) -> WhisperEncoder:
    """build_encoder.py:71-91 — q/k/v fused into one qkv weight with a ZERO K bias, conv weights unsqueezed to (1,3)."""
    p = "model.encoder."
    model.conv1.weight.value = _np(ckpt[p + "conv1.weight"].unsqueeze(2))
    model.conv1.bias.value = _np(ckpt[p + "conv1.bias"])
    model.conv2.weight.value = _np(ckpt[p + "conv2.weight"].unsqueeze(2))
    model.conv2.bias.value = _np(ckpt[p + "conv2.bias"])
    model.embed_positions_weight = _np(ckpt[p + "embed_positions.weight"].unsqueeze(0))
    for idx, layer in enumerate(model.layers):
        q = f"{p}layers.{idx}."
        layer.self_attn.qkv.weight.value = _np(torch.cat([ckpt[q + "self_attn.q_proj.weight"], ckpt[q + "self_attn.k_proj.weight"],
                                                          ckpt[q + "self_attn.v_proj.weight"]], dim=0))
        layer.self_attn.qkv.bias.value = _np(torch.cat([ckpt[q + "self_attn.q_proj.bias"],
                                                        torch.zeros_like(ckpt[q + "self_attn.q_proj.bias"]),
                                                        ckpt[q + "self_attn.v_proj.bias"]], dim=0))
        layer.self_attn.dense.weight.value = _np(ckpt[q + "self_attn.out_proj.weight"])
        layer.self_attn.dense.bias.value = _np(ckpt[q + "self_attn.out_proj.bias"])
        for ln in ("self_attn_layer_norm", "final_layer_norm"):
            getattr(layer, ln).weight.value = _np(ckpt[q + ln + ".weight"])
            getattr(layer, ln).bias.value = _np(ckpt[q + ln + ".bias"])
        for fc in ("fc1", "fc2"):
            getattr(layer, fc).weight.value = _np(ckpt[q + fc + ".weight"])
            getattr(layer, fc).bias.value = _np(ckpt[q + fc + ".bias"])
    model.layer_norm.weight.value = _np(ckpt[p + "layer_norm.weight"])
    model.layer_norm.bias.value = _np(ckpt[p + "layer_norm.bias"])
    return model


def load_decoder_from_hf(model: WhisperDecoder, ckpt: Dict[str, torch.Tensor]) -> WhisperDecoder:
    """build_decoder.py:71-101 (``out_proj`` -> ``dense``; proj_out.weight is the tied embedding)."""
    p = "model.decoder."
    model.embed_tokens.weight.value = _np(ckpt[p + "embed_tokens.weight"])
    model.embed_positions.weight.value = _np(ckpt[p + "embed_positions.weight"])
    for idx, layer in enumerate(model.layers):
        q = f"{p}layers.{idx}."
        for attn in ("self_attn", "encoder_attn"):
            a = getattr(layer, attn)
            a.q_proj.weight.value = _np(ckpt[f"{q}{attn}.q_proj.weight"])
            a.q_proj.bias.value = _np(ckpt[f"{q}{attn}.q_proj.bias"])
            a.k_proj.weight.value = _np(ckpt[f"{q}{attn}.k_proj.weight"])
            a.v_proj.weight.value = _np(ckpt[f"{q}{attn}.v_proj.weight"])
            a.v_proj.bias.value = _np(ckpt[f"{q}{attn}.v_proj.bias"])
            a.dense.weight.value = _np(ckpt[f"{q}{attn}.out_proj.weight"])
            a.dense.bias.value = _np(ckpt[f"{q}{attn}.out_proj.bias"])
        for ln in ("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm"):
            getattr(layer, ln).weight.value = _np(ckpt[q + ln + ".weight"])
            getattr(layer, ln).bias.value = _np(ckpt[q + ln + ".bias"])
        for fc in ("fc1", "fc2"):
            getattr(layer, fc).weight.value = _np(ckpt[q + fc + ".weight"])
            getattr(layer, fc).bias.value = _np(ckpt[q + fc + ".bias"])
    model.layer_norm.weight.value = _np(ckpt[p + "layer_norm.weight"])
    model.layer_norm.bias.value = _np(ckpt[p + "layer_norm.bias"])
    model.proj_out.weight.value = _np(ckpt["proj_out.weight"] if "proj_out.weight" in ckpt else ckpt[p + "embed_tokens.weight"])
    return model


def encoder_from_config(config: Dict, dtype=None) -> WhisperEncoder:
    """The constructor call of build_encoder.py:48-56."""
    return WhisperEncoder(d_model=config["d_model"], num_mel_bins=config["num_mel_bins"],
                          max_source_positions=config["max_source_positions"], encoder_layers=config["encoder_layers"],
                          encoder_attention_heads=config["encoder_attention_heads"],
                          activation_function=config["activation_function"], encoder_ffn_dim=config["encoder_ffn_dim"], dtype=dtype)


def decoder_from_config(config: Dict, dtype=None) -> WhisperDecoder:
    """The constructor call of build_decoder.py:45-56."""
    return WhisperDecoder(pad_token_id=config["pad_token_id"], max_target_positions=config["max_target_positions"],
                          max_source_positions=config["max_source_positions"], d_model=config["d_model"],
                          scale_embedding=config["scale_embedding"], vocab_size=config["vocab_size"],
                          decoder_layers=config["decoder_layers"], decoder_attention_heads=config["decoder_attention_heads"],
                          activation_function=config["activation_function"], decoder_ffn_dim=config["decoder_ffn_dim"], dtype=dtype)


def export_hf_state_dict(encoder: Optional[WhisperEncoder], decoder: Optional[WhisperDecoder]) -> Dict[str, torch.Tensor]:
    """Inverse of the binders: module parameters -> oracle state_dict keys.  This is what the
    native runtime (WhisperEngine / wb_model_load_tensor) consumes, so modules bound by the reference's own binder
    code can be handed to the fast path."""
    sd: Dict[str, torch.Tensor] = {}
    if encoder is not None:
        p = "model.encoder."
        d = encoder.d_model
        sd[p + "conv1.weight"] = encoder.conv1.weight.data.reshape(d, encoder.num_mel_bins, 3)
        sd[p + "conv1.bias"] = encoder.conv1.bias.data
        sd[p + "conv2.weight"] = encoder.conv2.weight.data.reshape(d, d, 3)
        sd[p + "conv2.bias"] = encoder.conv2.bias.data
        sd[p + "embed_positions.weight"] = encoder._positions(encoder.conv1.weight.device).reshape(-1, d)
        for idx, layer in enumerate(encoder.layers):
            q = f"{p}layers.{idx}."
            w, b = layer.self_attn.qkv.weight.data, layer.self_attn.qkv.bias.data
            sd[q + "self_attn.q_proj.weight"], sd[q + "self_attn.k_proj.weight"], sd[q + "self_attn.v_proj.weight"] = w[:d], w[d:2 * d], w[2 * d:]
            sd[q + "self_attn.q_proj.bias"], sd[q + "self_attn.v_proj.bias"] = b[:d], b[2 * d:]
            sd[q + "self_attn.out_proj.weight"] = layer.self_attn.dense.weight.data
            sd[q + "self_attn.out_proj.bias"] = layer.self_attn.dense.bias.data
            for name in ("self_attn_layer_norm", "final_layer_norm", "fc1", "fc2"):
                sd[q + name + ".weight"] = getattr(layer, name).weight.data
                sd[q + name + ".bias"] = getattr(layer, name).bias.data
        sd[p + "layer_norm.weight"] = encoder.layer_norm.weight.data
        sd[p + "layer_norm.bias"] = encoder.layer_norm.bias.data
    if decoder is not None:
        p = "model.decoder."
        sd[p + "embed_tokens.weight"] = decoder.embed_tokens.weight.data
        sd[p + "embed_positions.weight"] = decoder.embed_positions.weight.data
        for idx, layer in enumerate(decoder.layers):
            q = f"{p}layers.{idx}."
            for attn in ("self_attn", "encoder_attn"):
                a = getattr(layer, attn)
                sd[f"{q}{attn}.q_proj.weight"], sd[f"{q}{attn}.q_proj.bias"] = a.q_proj.weight.data, a.q_proj.bias.data
                sd[f"{q}{attn}.k_proj.weight"] = a.k_proj.weight.data
                sd[f"{q}{attn}.v_proj.weight"], sd[f"{q}{attn}.v_proj.bias"] = a.v_proj.weight.data, a.v_proj.bias.data
                sd[f"{q}{attn}.out_proj.weight"], sd[f"{q}{attn}.out_proj.bias"] = a.dense.weight.data, a.dense.bias.data
            for name in ("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm", "fc1", "fc2"):
                sd[q + name + ".weight"] = getattr(layer, name).weight.data
                sd[q + name + ".bias"] = getattr(layer, name).bias.data
        sd[p + "layer_norm.weight"] = decoder.layer_norm.weight.data
        sd[p + "layer_norm.bias"] = decoder.layer_norm.bias.data
        sd["proj_out.weight"] = sd[p + "embed_tokens.weight"]
        if not torch.equal(decoder.proj_out.weight.data, decoder.embed_tokens.weight.data):
            raise ValueError("proj_out.weight must be tied to embed_tokens.weight (modeling_whisper.py:1335)")
    return sd
