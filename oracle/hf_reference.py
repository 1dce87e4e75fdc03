"""Import the REAL reference (the simplified HF Whisper vendored in /root/reference) — build container only.

TEST INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box, so nothing that runs there
imports this module; it is used by oracle/make_golden.py and by ``tests/test_oracle_vs_reference.py``
(auto-skipped when the tree is absent).  Recipe: SURVEY.md Appendix D.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
from typing import Dict

import torch

REF_SRC = "/root/reference/transformers/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "transformers", "models", "whisper"))


def import_reference():
    """Returns (WhisperConfig, WhisperForConditionalGeneration, BaseModelOutput) of the vendored tree."""
    if not available():
        raise RuntimeError("reference tree not mounted")
    sys.dont_write_bytecode = True  # /root/reference is read-only
    loaded = sys.modules.get("transformers")
    if loaded is not None and not getattr(loaded, "__file__", "").startswith(REF_SRC):
        raise RuntimeError("a different `transformers` is already imported; run the oracle in a fresh process")
    stub = types.ModuleType("transformers.dependency_versions_check")
    stub.dep_version_check = lambda *a, **k: None  # bypass tokenizers<0.14 pin (dependency_versions_check.py:57)
    sys.modules["transformers.dependency_versions_check"] = stub
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from transformers.models.whisper.configuration_whisper import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperForConditionalGeneration
    from transformers.modeling_outputs import BaseModelOutput
    return WhisperConfig, WhisperForConditionalGeneration, BaseModelOutput


def build_reference_model(cfg: Dict, sd: Dict[str, torch.Tensor]):
    """Instantiate the reference model for ``cfg`` (oracle/synth.make_config) and load ``sd`` into it."""
    WhisperConfig, WhisperForConditionalGeneration, _ = import_reference()
    hf_cfg = WhisperConfig(
        vocab_size=cfg["vocab_size"], num_mel_bins=cfg["num_mel_bins"], d_model=cfg["d_model"],
        encoder_layers=cfg["encoder_layers"], decoder_layers=cfg["decoder_layers"],
        encoder_attention_heads=cfg["encoder_attention_heads"],
        decoder_attention_heads=cfg["decoder_attention_heads"],
        encoder_ffn_dim=cfg["encoder_ffn_dim"], decoder_ffn_dim=cfg["decoder_ffn_dim"],
        max_source_positions=cfg["max_source_positions"], max_target_positions=cfg["max_target_positions"],
        pad_token_id=cfg["pad_token_id"], bos_token_id=cfg["bos_token_id"], eos_token_id=cfg["eos_token_id"],
        decoder_start_token_id=cfg["decoder_start_token_id"],
        suppress_tokens=list(cfg["suppress_tokens"]), begin_suppress_tokens=list(cfg["begin_suppress_tokens"]),
        max_length=cfg["max_length"],
    )
    hf_cfg.forced_decoder_ids = [list(p) for p in cfg["forced_decoder_ids"]]
    with contextlib.redirect_stdout(io.StringIO()):
        model = WhisperForConditionalGeneration(hf_cfg)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k == "proj_out.weight" for k in missing), missing
    model.tie_weights()
    model.eval()
    # a model built from a bare config gets generation_config.max_length=20 / suppress=None
    gc = model.generation_config
    gc.max_length = cfg["max_length"]
    gc.suppress_tokens = list(cfg["suppress_tokens"])
    gc.begin_suppress_tokens = list(cfg["begin_suppress_tokens"])
    gc.forced_decoder_ids = [list(p) for p in cfg["forced_decoder_ids"]]
    gc.pad_token_id = cfg["pad_token_id"]
    gc.eos_token_id = cfg["eos_token_id"]
    gc.decoder_start_token_id = cfg["decoder_start_token_id"]
    return model


@torch.no_grad()
def reference_generate(model, mel: torch.Tensor, max_length: int = None) -> torch.Tensor:
    """``hf_model.generate(input_features)`` exactly as run.py:308 calls it (debug prints swallowed)."""
    if max_length is not None:
        model.generation_config.max_length = max_length
    with contextlib.redirect_stdout(io.StringIO()):
        return model.generate(mel)


@torch.no_grad()
def reference_encode(model, mel: torch.Tensor) -> torch.Tensor:
    return model.model.encoder(mel).last_hidden_state


@torch.no_grad()
def reference_step_logits(model, enc: torch.Tensor, ids: torch.Tensor, steps: int):
    """Teacher-forced per-step logits through the reference's forward with its own past_key_values."""
    _, _, BaseModelOutput = import_reference()
    out = []
    past = None
    for t in range(steps):
        o = model(decoder_input_ids=ids[:, t:t + 1] if past is not None else ids[:, :1],
                  encoder_outputs=BaseModelOutput(last_hidden_state=enc), past_key_values=past, use_cache=True)
        past = o.past_key_values
        out.append(o.logits[:, -1, :].clone())
    return out, past
