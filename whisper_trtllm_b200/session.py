"""``Session`` / ``TensorInfo`` with the interface of the reference's ``tensorrt_llm.runtime`` (runtime/session.py:28-178):

    session = Session.from_serialized_engine(engine_buffer)
    outputs_info = session.infer_shapes([TensorInfo(name, dtype, shape), ...])
    ok = session.run(inputs, outputs, stream)        # async enqueue on a raw cudaStream_t; True = enqueued

plus the "engine" files the reference's ``build_encoder.py`` / ``build_decoder.py`` produce.  A serialized engine
here is not a TensorRT plan: it is the model's constructor kwargs + its weights (nothing is traced or compiled,
the kernels are already in libwhisper_b200.so), in a small self-describing container:

    b"WB200ENG" | u32 version | u64 header_len | header JSON | raw little-endian fp32 tensors

Tensor contract (model.py:113-124, 464-516), batch generalised from 1 to B:
  encoder   in  data f32 [B,80,3000], length f32 [B] (unused)              out hidden_states f32 [B,1500,d]
  decoder   in  data i32 [B,1], length i32 [B], encoder_hidden_states f32 [B,1500,d],
                self_past_key/value f32 [L,H,T,64] (B == 1) or [L,B,H,T,64],
                cross_past_key/value f32 [L,H,1500,64] or [L,B,H,1500,64],
                past_self_cache_mask f32 [n+1], past_cross_cache_mask f32 [m+1]   (lengths live in the SHAPES)
            out hidden_states (= logits) f32 [B,1,V], next_self_keys/values [..., n'+1, 64], next_cross_keys/values
"""
from __future__ import annotations

import io
import json
import struct
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import _abi
from .model import WhisperDecoder, WhisperEncoder

MAGIC = b"WB200ENG"
ENGINE_VERSION = 1


class DataType:
    """Stand-ins for ``trt.float32`` / ``trt.int32`` ... in TensorInfo (run.py:21-33 maps them to torch dtypes)."""
    float32 = torch.float32
    float16 = torch.float16
    bfloat16 = torch.bfloat16
    int32 = torch.int32
    int8 = torch.int8


@dataclass
class TensorInfo:
    name: str
    dtype: Any
    shape: tuple


# ------------------------------------------------------------------------------------------------------------
# engine container
# ------------------------------------------------------------------------------------------------------------
def _ctor_kwargs(model) -> Dict:
    if isinstance(model, WhisperEncoder):
        layer = model.layers[0]
        return dict(d_model=model.d_model, num_mel_bins=model.num_mel_bins, max_source_positions=model.max_source_positions,
                    encoder_layers=len(model.layers), encoder_attention_heads=layer.self_attn.num_attention_heads,
                    activation_function="gelu", encoder_ffn_dim=layer.fc1.out_features)
    layer = model.layers[0]
    return dict(pad_token_id=model.padding_idx, max_target_positions=model.max_target_positions,
                max_source_positions=model.max_source_positions, d_model=model.d_model,
                scale_embedding=model.embed_scale != 1.0, vocab_size=model.vocab_size, decoder_layers=model.decoder_layers,
                decoder_attention_heads=model.decoder_attention_heads, activation_function="gelu",
                decoder_ffn_dim=layer.fc1.out_features)


def serialize_engine(model, path: Optional[str] = None, precision: str = "float32") -> bytes:
    """``builder.build_engine`` + ``serialize_engine`` of build_encoder.py:104-109: module -> engine bytes
    (written to ``path`` when given).  ``precision`` is the engine's compute dtype ('float32' | 'bfloat16')."""
    kind = "WhisperEncoder" if isinstance(model, WhisperEncoder) else "WhisperDecoder"
    if not isinstance(model, (WhisperEncoder, WhisperDecoder)):
        raise TypeError("serialize_engine expects a WhisperEncoder or WhisperDecoder module")
    tensors = {name: p.data.detach().to("cpu", torch.float32).contiguous().numpy() for name, p in model.named_parameters()}
    if kind == "WhisperEncoder":
        tensors["embed_positions_weight"] = np.ascontiguousarray(np.asarray(model.embed_positions_weight, dtype=np.float32))
    index, off = [], 0
    for name, a in tensors.items():
        index.append({"name": name, "shape": list(a.shape), "offset": off})
        off += a.nbytes
    header = json.dumps({"kind": kind, "precision": precision, "kwargs": _ctor_kwargs(model), "tensors": index}).encode()
    buf = io.BytesIO()
    buf.write(MAGIC)
    buf.write(struct.pack("<IQ", ENGINE_VERSION, len(header)))
    buf.write(header)
    for a in tensors.values():
        buf.write(a.astype("<f4", copy=False).tobytes())
    data = buf.getvalue()
    if path is not None:
        with open(path, "wb") as f:
            f.write(data)
    return data


def deserialize_engine(engine_buffer: bytes):
    """engine bytes -> (kind, module with weights bound).  Fails loudly on a foreign / truncated buffer."""
    mv = memoryview(engine_buffer)
    if len(mv) < 20 or bytes(mv[:8]) != MAGIC:
        raise _abi.WhisperB200Error(-1, "not a whisper_b200 engine (bad magic) — TensorRT plans are not accepted")
    version, hlen = struct.unpack("<IQ", mv[8:20])
    if version != ENGINE_VERSION:
        raise _abi.WhisperB200Error(-1, f"engine version {version} not supported (expected {ENGINE_VERSION})")
    header = json.loads(bytes(mv[20:20 + hlen]))
    base = 20 + hlen
    cls = WhisperEncoder if header["kind"] == "WhisperEncoder" else WhisperDecoder
    model = cls(dtype=header["precision"], **header["kwargs"])
    params = dict(model.named_parameters())
    for t in header["tensors"]:
        n = int(np.prod(t["shape"])) if t["shape"] else 1
        start = base + t["offset"]
        if start + 4 * n > len(mv):
            raise _abi.WhisperB200Error(-1, f"engine truncated at tensor {t['name']}")
        a = np.frombuffer(mv[start:start + 4 * n], dtype="<f4").reshape(t["shape"])
        if t["name"] == "embed_positions_weight":
            model.embed_positions_weight = a
        else:
            params[t["name"]].value = a
    return header["kind"], model


# ------------------------------------------------------------------------------------------------------------
# Session
# ------------------------------------------------------------------------------------------------------------
class Session(object):
    """runtime/session.py:35-178.  Use ``Session.from_serialized_engine`` (or ``Session.from_model``)."""

    ENCODER_INPUTS = ("data", "length")
    DECODER_INPUTS = ("data", "length", "encoder_hidden_states", "self_past_key", "self_past_value", "cross_past_key",
                      "cross_past_value", "past_self_cache_mask", "past_cross_cache_mask")

    def __init__(self, **kwargs):
        self._model = None
        self._kind = None
        self._shapes: Dict[str, tuple] = {}
        self._out_info: List[TensorInfo] = []

    def _init(self, kind, model):
        self._kind, self._model = kind, model
        return self

    @staticmethod
    def from_serialized_engine(engine) -> "Session":
        kind, model = deserialize_engine(engine)
        return Session()._init(kind, model)

    @staticmethod
    def from_model(model) -> "Session":
        kind = "WhisperEncoder" if isinstance(model, WhisperEncoder) else "WhisperDecoder"
        return Session()._init(kind, model)

    @property
    def engine(self):
        return self._model

    # ---- shapes ------------------------------------------------------------------------------------------
    def infer_shapes(self, inputs: List[TensorInfo], context=None) -> Optional[List[TensorInfo]]:
        """Set the input shapes and return the output TensorInfos (None on a wrong name/dtype, like the reference,
        session.py:116-146)."""
        m = self._model
        known = self.ENCODER_INPUTS if self._kind == "WhisperEncoder" else self.DECODER_INPUTS
        shapes = {}
        for i in inputs:
            if i.name not in known:
                return None
            want_int = self._kind == "WhisperDecoder" and i.name in ("data", "length")
            if (i.dtype == torch.int32) != want_int:
                return None
            shapes[i.name] = tuple(int(s) for s in i.shape)
        self._shapes = shapes
        f32 = DataType.float32
        if self._kind == "WhisperEncoder":
            B = shapes["data"][0]
            out = [TensorInfo("hidden_states", f32, (B, m.max_source_positions, m.d_model))]
        else:
            B = shapes["data"][0]
            sk = shapes["self_past_key"]
            n_self = min(shapes["past_self_cache_mask"][0] - 1, sk[-2])
            m_cross = shapes["past_cross_cache_mask"][0] - 1
            S = m.max_source_positions
            if not 0 <= m_cross <= S:
                return None
            lead = sk[:-2]
            out = [TensorInfo("hidden_states", f32, (B, shapes["data"][1], m.vocab_size)),
                   TensorInfo("next_self_keys", f32, lead + (n_self + 1, m.d_head)),
                   TensorInfo("next_self_values", f32, lead + (n_self + 1, m.d_head)),
                   TensorInfo("next_cross_keys", f32, lead + (S, m.d_head)),
                   TensorInfo("next_cross_values", f32, lead + (S, m.d_head))]
        self._out_info = out
        return out

    # ---- run ---------------------------------------------------------------------------------------------
    def _wrap(self, name: str, value, dtype, shape, device) -> torch.Tensor:
        """Tensor-or-raw-pointer -> tensor (a raw pointer is viewed with the shape given to infer_shapes)."""
        if isinstance(value, torch.Tensor):
            return value
        if shape is None:
            raise _abi.WhisperB200Error(-1, f"{name} was passed as a raw pointer: call infer_shapes first")
        n = int(np.prod(shape)) if len(shape) else 1
        esz = torch.empty((), dtype=dtype).element_size()
        mem = _RawDeviceMemory(int(value), n * esz, device)
        return torch.as_tensor(mem, device=device).view(dtype).view(*shape)

    def run(self, inputs: Dict[str, Any], outputs: Dict[str, Any], stream, context=None) -> bool:
        """Enqueue on ``stream`` (raw cudaStream_t integer, or a torch stream); returns True once enqueued.
        The caller synchronises (run.py:85-87)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        ext = stream if isinstance(stream, torch.cuda.Stream) else torch.cuda.ExternalStream(int(stream or 0), device=dev)
        info = {t.name: t for t in self._out_info}
        with torch.cuda.stream(ext):
            if self._kind == "WhisperEncoder":
                mel = self._wrap("data", inputs["data"], torch.float32, self._shapes.get("data"), dev)
                res = {"hidden_states": self._model(mel)}
            else:
                def get(name, dtype=torch.float32):
                    return self._wrap(name, inputs[name], dtype, self._shapes.get(name), dev)
                logits, nsk, nsv, nck, ncv = self._model(
                    get("data", torch.int32), get("encoder_hidden_states"), get("self_past_key"), get("self_past_value"),
                    get("cross_past_key"), get("cross_past_value"), get("past_self_cache_mask"), get("past_cross_cache_mask"))
                res = {"hidden_states": logits, "next_self_keys": nsk, "next_self_values": nsv, "next_cross_keys": nck,
                       "next_cross_values": ncv}
            for name, dst in outputs.items():
                src = res[name]
                t = info.get(name)
                dst_t = self._wrap(name, dst, torch.float32, tuple(t.shape) if t is not None else tuple(src.shape), dev)
                if tuple(dst_t.shape) != tuple(src.shape):
                    raise _abi.WhisperB200Error(-1, f"output {name}: buffer shape {tuple(dst_t.shape)} != result {tuple(src.shape)}")
                if dst_t.data_ptr() != src.data_ptr():
                    dst_t.copy_(src)      # engine outputs are fp32 (mark_output(..., float32), model.py:464-468)
        return True

    def _debug_run(self, inputs: Dict[str, torch.Tensor], context=None) -> Dict[str, torch.Tensor]:
        """Synchronous run with freshly allocated outputs (session.py:180-205)."""
        infos = self.infer_shapes([TensorInfo(n, t.dtype, tuple(t.shape)) for n, t in inputs.items()])
        outputs = {t.name: torch.empty(tuple(t.shape), dtype=torch.float32, device="cuda") for t in infos}
        stream = torch.cuda.current_stream()
        self.run(inputs, outputs, stream.cuda_stream)
        stream.synchronize()
        return outputs


class _RawDeviceMemory:
    """__cuda_array_interface__ wrapper so that torch can view caller-owned device memory given as an address."""

    def __init__(self, address: int, nbytes: int, device):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (address, False), "version": 3,
                                         "strides": None}
