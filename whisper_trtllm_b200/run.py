"""The greedy-search session interface of the reference's ``examples/whisper/run.py``: runner classes
``WhisperEncoder`` / ``WhisperDecoder`` (run.py:57-148), ``get_logits_processor`` (:150-162),
``get_stopping_criteria`` (:164-169) and ``greedy_search`` (:171-227), with the same call signatures.

Two execution paths behind ``greedy_search``:
  * FAST (default when it applies): the model is this module's ``WhisperDecoder`` runner, the processors are the three
    standard ones and the criterion is ``MaxLengthCriteria`` -> the whole loop runs on the device inside
    libwhisper_b200 (paged in-place KV cache, fused suppress/argmax/EOS bookkeeping, no per-step host sync).
  * GENERIC: any ``model(last_ids, encoder_outputs, past) -> (logits, past)`` callable and arbitrary processor /
    criteria callables, stepping exactly like run.py:195-225.
Both produce the ids the reference's loop produces (tests/test_gpu_dropin.py).
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, List, Optional

import torch

from . import _abi
from ._abi import ptr, stream_handle
from .engine import WhisperEngine
from .model import (WhisperDecoder as WhisperDecoderModule, WhisperEncoder as WhisperEncoderModule, decoder_from_config,
                    encoder_from_config, export_hf_state_dict, load_decoder_from_hf, load_encoder_from_hf)
from .session import DataType, Session, TensorInfo, serialize_engine

ENCODER_ENGINE, DECODER_ENGINE, CONFIG_PKL = "WhisperEncoder.engine", "WhisperDecoder.engine", "config.pkl"


# ------------------------------------------------------------------------------------------------------------
# logits processors / stopping criteria (transformers/generation/logits_process.py:1281-1328, stopping_criteria.py:61-70)
# ------------------------------------------------------------------------------------------------------------
class LogitsProcessorList(list):
    def __call__(self, input_ids, scores):
        for p in self:
            scores = p(input_ids, scores)
        return scores


class SuppressTokensLogitsProcessor:
    """scores[:, suppress_tokens] = -inf at every step (logits_process.py:1300-1310)."""

    def __init__(self, suppress_tokens):
        self.suppress_tokens = list(suppress_tokens)

    def __call__(self, input_ids, scores):
        scores[:, self.suppress_tokens] = -float("inf")
        return scores


class SuppressTokensAtBeginLogitsProcessor:
    """scores[:, begin_suppress_tokens] = -inf when len(input_ids) == begin_index (logits_process.py:1281-1297)."""

    def __init__(self, begin_suppress_tokens, begin_index):
        self.begin_suppress_tokens = list(begin_suppress_tokens)
        self.begin_index = begin_index

    def __call__(self, input_ids, scores):
        if input_ids.shape[1] == self.begin_index:
            scores[:, self.begin_suppress_tokens] = -float("inf")
        return scores


class ForceTokensLogitsProcessor:
    """When len(input_ids) is a key of the map: everything -inf, forced token 0 (logits_process.py:1313-1328)."""

    def __init__(self, force_token_map):
        self.force_token_map = dict(force_token_map)

    def __call__(self, input_ids, scores):
        generation_idx = input_ids.shape[-1]
        current_token = self.force_token_map.get(generation_idx, None)
        if current_token is not None:
            scores[:, :] = -float("inf")
            scores[:, current_token] = 0
        return scores


class MaxLengthCriteria:
    """stopping_criteria.py:61-70: stop when len(input_ids) >= max_length."""

    def __init__(self, max_length: int, max_position_embeddings: Optional[int] = None):
        self.max_length = max_length
        self.max_position_embeddings = max_position_embeddings

    def __call__(self, input_ids, scores, **kwargs) -> bool:
        return input_ids.shape[-1] >= self.max_length


class StoppingCriteriaList(list):
    def __call__(self, input_ids, scores, **kwargs) -> bool:
        return any(c(input_ids, scores) for c in self)

    @property
    def max_length(self) -> Optional[int]:
        for c in self:
            if isinstance(c, MaxLengthCriteria):
                return c.max_length
        return None


def get_logits_processor(config, input_ids_seq_length):
    """run.py:150-162, same order: suppress -> begin-suppress -> force."""
    processors = LogitsProcessorList()
    processors.append(SuppressTokensLogitsProcessor(config["suppress_tokens"]))
    begin_index = input_ids_seq_length
    begin_index = begin_index if config["forced_bos_token_id"] is None else begin_index + 1
    begin_index += config["forced_decoder_ids"][-1][0]
    processors.append(SuppressTokensAtBeginLogitsProcessor(config["begin_suppress_tokens"], begin_index))
    processors.append(ForceTokensLogitsProcessor(config["forced_decoder_ids"]))
    return processors


def get_stopping_criteria(config):
    """run.py:164-169."""
    criteria = StoppingCriteriaList()
    criteria.append(MaxLengthCriteria(max_length=config["max_length"], max_position_embeddings=None))
    return criteria


# ------------------------------------------------------------------------------------------------------------
# engine build (build_encoder.py / build_decoder.py): HF state_dict -> engine files + config.pkl
# ------------------------------------------------------------------------------------------------------------
def build_encoder(config: Dict, ckpt: Dict[str, torch.Tensor], engine_dir: Optional[str] = None,
                  engine_precision: str = "float32") -> bytes:
    """build_encoder.py:30-109: construct the module with the reference's ctor call, bind the HF weights with the
    reference's mapping, write ``WhisperEncoder.engine`` (+ ``config.pkl``, :42-45)."""
    model = load_encoder_from_hf(encoder_from_config(config, dtype=engine_precision), ckpt)
    if engine_dir is not None:
        os.makedirs(engine_dir, exist_ok=True)
        with open(os.path.join(engine_dir, CONFIG_PKL), "wb") as f:
            pickle.dump(dict(config), f)
    return serialize_engine(model, os.path.join(engine_dir, ENCODER_ENGINE) if engine_dir else None, engine_precision)


def build_decoder(config: Dict, ckpt: Dict[str, torch.Tensor], engine_dir: Optional[str] = None,
                  engine_precision: str = "float32") -> bytes:
    """build_decoder.py:30-119."""
    model = load_decoder_from_hf(decoder_from_config(config, dtype=engine_precision), ckpt)
    if engine_dir is not None:
        os.makedirs(engine_dir, exist_ok=True)
    return serialize_engine(model, os.path.join(engine_dir, DECODER_ENGINE) if engine_dir else None, engine_precision)


def _engine_bytes(args, engine, name) -> bytes:
    if engine is not None:
        return engine
    with open(os.path.join(args.engine_dir, name), "rb") as f:
        return f.read()


# ------------------------------------------------------------------------------------------------------------
# runner classes (run.py:57-148)
# ------------------------------------------------------------------------------------------------------------
class WhisperEncoder:
    """run.py:57-91: ``WhisperEncoder(args, config)(input) -> hidden_states``; batch generalised to B."""

    def __init__(self, args=None, config=None, engine: Optional[bytes] = None):
        self.config = config
        self.session = Session.from_serialized_engine(_engine_bytes(args, engine, ENCODER_ENGINE))
        self.fast = None          # set by link(): the shared native engine
        self._last = None

    def __call__(self, input):
        if self.fast is not None:
            hidden = self.fast.engine.encode(input.contiguous().float(), return_hidden=True)
            torch.cuda.current_stream().synchronize()
            self.fast.note_encoder_output(hidden)
            return hidden
        inputs = {"data": input.contiguous(), "length": torch.ones(input.shape[0], device=input.device)}
        infos = self.session.infer_shapes([TensorInfo("data", DataType.float32, tuple(input.shape)),
                                           TensorInfo("length", DataType.float32, (input.shape[0],))])
        outputs = {o.name: torch.zeros(*o.shape, dtype=torch.float32, device=input.device) for o in infos}
        stream = torch.cuda.current_stream()
        ok = self.session.run(inputs, outputs, stream.cuda_stream)
        stream.synchronize()
        assert ok
        return outputs["hidden_states"]


class WhisperDecoder:
    """run.py:93-148: ``WhisperDecoder(args, config)(decoder_input_ids, encoder_outputs, past_key_values)
    -> (logits, (next_self_keys, next_self_values, next_cross_keys, next_cross_values))``.

    (The reference reads the module-global ``config`` in ``__call__`` — a bug, run.py:110-112; here ``self.config``.)"""

    def __init__(self, args=None, config=None, engine: Optional[bytes] = None):
        self.config = config
        self.session = Session.from_serialized_engine(_engine_bytes(args, engine, DECODER_ENGINE))
        self.fast = None

    def __call__(self, decoder_input_ids, encoder_outputs, past_key_values):
        config = self.config
        dev = encoder_outputs.device
        B = decoder_input_ids.shape[0]
        L, H = config["decoder_layers"], config["decoder_attention_heads"]
        dh = config["d_model"] // H
        S = config["max_source_positions"]
        lead = (L, H) if B == 1 else (L, B, H)
        inputs = {
            "data": decoder_input_ids.to(dtype=torch.int32, device=dev),
            "length": torch.ones(B, dtype=torch.int32, device=dev),
            "encoder_hidden_states": encoder_outputs.to(dtype=torch.float32),
        }
        if past_key_values is None:       # step 0: dummy caches, masks of length 1 (cache length 0), run.py:108-119
            inputs["self_past_key"] = torch.zeros(*lead, 1, dh, device=dev)
            inputs["self_past_value"] = torch.zeros(*lead, 1, dh, device=dev)
            inputs["cross_past_key"] = torch.zeros(*lead, S, dh, device=dev)
            inputs["cross_past_value"] = torch.zeros(*lead, S, dh, device=dev)
            inputs["past_self_cache_mask"] = torch.zeros(1, device=dev)
            inputs["past_cross_cache_mask"] = torch.zeros(1, device=dev)
        else:
            for name, t in zip(("self_past_key", "self_past_value", "cross_past_key", "cross_past_value"), past_key_values):
                inputs[name] = t.to(dtype=torch.float32, device=dev)
            inputs["past_self_cache_mask"] = torch.zeros(int(1 + past_key_values[0].shape[-2]), device=dev)
            inputs["past_cross_cache_mask"] = torch.zeros(int(1 + S), device=dev)
        infos = self.session.infer_shapes([TensorInfo(k, v.dtype, tuple(v.shape)) for k, v in inputs.items()])
        outputs = {}
        for o in infos:
            # the cross caches are returned as the caller's own storage when they are reused (no re-copy)
            if past_key_values is not None and o.name in ("next_cross_keys", "next_cross_values"):
                outputs[o.name] = inputs["cross_past_key" if o.name.endswith("keys") else "cross_past_value"]
            else:
                outputs[o.name] = torch.empty(*o.shape, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream()
        ok = self.session.run(inputs, outputs, stream.cuda_stream)
        stream.synchronize()
        assert ok
        return outputs["hidden_states"], (outputs["next_self_keys"], outputs["next_self_values"],
                                          outputs["next_cross_keys"], outputs["next_cross_values"])


class FastPath:
    """The native engine shared by a runner encoder/decoder pair (``link``)."""

    def __init__(self, engine: WhisperEngine):
        self.engine = engine
        self._enc_ref = None      # STRONG reference to the tensor whose cross K/V the engine holds: identity, not address
        self._enc_version = -1

    def note_encoder_output(self, hidden: torch.Tensor):
        # holding the tensor pins its storage: the caching allocator cannot hand the same address to another tensor, and
        # `is` + the version counter cannot be fooled by a different tensor of the same shape
        self._enc_ref = hidden
        self._enc_version = hidden._version

    def has_encoder_output(self, hidden: torch.Tensor) -> bool:
        return self._enc_ref is hidden and self._enc_version == hidden._version


def link(encoder: WhisperEncoder, decoder: WhisperDecoder, config: Dict, max_batch: int = 1, dtype: Optional[str] = None,
         enc_chunk: Optional[int] = None) -> FastPath:
    """Pack the two engines' weights into ONE native runtime (wb_model + wb_session) so that ``greedy_search`` can run
    the whole loop on the device.  Without ``link`` everything still works through the per-step Session path."""
    enc_m, dec_m = encoder.session.engine, decoder.session.engine
    dt = dtype or {torch.float32: "float32", torch.bfloat16: "bfloat16"}[dec_m.dtype]
    cfg = dict(config)
    cfg["max_length"] = cfg["max_target_positions"]   # the stopping criterion's max_length is applied per call
    eng = WhisperEngine(cfg, export_hf_state_dict(enc_m, dec_m), dtype=dt, max_batch=max_batch, enc_chunk=enc_chunk,
                        device=dec_m.embed_tokens.weight.device)
    fast = FastPath(eng)
    encoder.fast = fast
    decoder.fast = fast
    return fast


# ------------------------------------------------------------------------------------------------------------
# greedy search (run.py:171-227 == generation/utils.py:1474-1529)
# ------------------------------------------------------------------------------------------------------------
def _standard_processors(logits_processor, stopping_criteria):
    """-> (suppress, begin_suppress, begin_index, force_map, max_length) when the lists are exactly the three
    standard processors in the reference's order + a MaxLengthCriteria; else None."""
    if logits_processor is None or stopping_criteria is None or len(logits_processor) != 3 or len(stopping_criteria) != 1:
        return None
    a, b, c = logits_processor
    if not (type(a) is SuppressTokensLogitsProcessor and type(b) is SuppressTokensAtBeginLogitsProcessor
            and type(c) is ForceTokensLogitsProcessor and type(stopping_criteria[0]) is MaxLengthCriteria):
        return None
    return a.suppress_tokens, b.begin_suppress_tokens, b.begin_index, c.force_token_map, stopping_criteria[0].max_length


def greedy_search(model, encoder_outputs, input_ids: torch.Tensor, logits_processor=None, stopping_criteria=None,
                  pad_token_id=None, eos_token_id=None):
    fast = getattr(model, "fast", None)
    std = _standard_processors(logits_processor, stopping_criteria)
    if (fast is not None and std is not None and input_ids.shape[1] == 1 and input_ids.shape[0] <= fast.engine.max_batch
            and pad_token_id == fast.engine.config["pad_token_id"] and eos_token_id == fast.engine.config["eos_token_id"]
            and bool((input_ids == fast.engine.config["decoder_start_token_id"]).all())
            and 2 <= std[4] <= fast.engine.config["max_target_positions"]):
        return _greedy_search_device(fast, encoder_outputs, input_ids, std)
    return _greedy_search_generic(model, encoder_outputs, input_ids, logits_processor, stopping_criteria, pad_token_id, eos_token_id)


def _greedy_search_device(fast: FastPath, encoder_outputs, input_ids, std):
    suppress, begin_suppress, begin_index, force_map, max_length = std
    eng = fast.engine
    B = input_ids.shape[0]
    eng.set_generation(suppress, begin_suppress, begin_index, sorted(force_map.items()))
    if not fast.has_encoder_output(encoder_outputs):       # encoder states produced elsewhere: project the cross K/V
        eng.set_encoder_output(encoder_outputs.to(eng.device).contiguous())
        fast.note_encoder_output(encoder_outputs)
    ids = eng.greedy(B, max_new_tokens=max_length - 1)
    return ids.to(device=input_ids.device, dtype=input_ids.dtype)


def _greedy_search_generic(model, encoder_outputs, input_ids, logits_processor, stopping_criteria, pad_token_id, eos_token_id):
    """Step-by-step loop with the reference's exact order of operations (run.py:183-227); argmax runs in the
    library's warp-shuffle kernel, the processors are the caller's callables."""
    eos_token_id = [eos_token_id] if isinstance(eos_token_id, int) else list(eos_token_id)
    eos_token_id_tensor = torch.tensor(eos_token_id).to(input_ids.device)
    scores = None
    unfinished_sequences = torch.ones(input_ids.shape[0], dtype=torch.int32, device=input_ids.device)
    past_key_values = None
    while True:
        output, past_key_values = model(input_ids[:, -1:], encoder_outputs, past_key_values)
        next_token_logits = output[:, -1, :]
        next_tokens_scores = logits_processor(input_ids, next_token_logits)
        next_tokens = _argmax(next_tokens_scores).to(input_ids.device)
        next_tokens = next_tokens * unfinished_sequences + pad_token_id * (1 - unfinished_sequences)
        input_ids = torch.cat([input_ids, next_tokens[:, None].to(input_ids.dtype)], dim=-1)
        unfinished_sequences = unfinished_sequences.mul(
            next_tokens.tile(eos_token_id_tensor.shape[0], 1).ne(eos_token_id_tensor.unsqueeze(1)).prod(dim=0).to(torch.int32))
        this_peer_finished = bool(unfinished_sequences.max() == 0)
        if stopping_criteria(input_ids, scores):
            this_peer_finished = True
        if this_peer_finished:
            break
    return input_ids


def _argmax(scores: torch.Tensor) -> torch.Tensor:
    """First maximal index per row (torch.argmax semantics) through wb_argmax."""
    if not scores.is_cuda:
        raise _abi.WhisperB200Error(-101, "scores must be a CUDA tensor: libwhisper_b200 has no CPU path")
    s = scores.float().contiguous() if scores.dtype != torch.float32 or not scores.is_contiguous() else scores
    if s.shape[1] % 4 != 0:
        raise _abi.WhisperB200Error(-1, "vocab size must be a multiple of 4")
    out = torch.empty(s.shape[0], dtype=torch.int32, device=s.device)
    _abi.call("wb_argmax", ptr(s), s.stride(0), s.shape[0], s.shape[1], None, 0, ptr(out), stream_handle())
    return out


def transcribe(encoder: WhisperEncoder, decoder: WhisperDecoder, config: Dict, input_features: torch.Tensor) -> torch.Tensor:
    """One iteration of the reference's main loop (run.py:270-284) for a batch of log-mel windows."""
    encoder_outputs = encoder(input_features)
    B = input_features.shape[0]
    input_ids = torch.full((B, 1), config["decoder_start_token_id"], dtype=torch.int32, device=input_features.device)
    return greedy_search(model=decoder, encoder_outputs=encoder_outputs, input_ids=input_ids,
                         logits_processor=get_logits_processor(config, input_ids.shape[-1]),
                         stopping_criteria=get_stopping_criteria(config),
                         pad_token_id=config["pad_token_id"], eos_token_id=config["eos_token_id"])
