"""HF checkpoint directory -> (config dict, state_dict) for the engines.

The reference builds its engines from ``WhisperForConditionalGeneration.from_pretrained(args.whisper)``:
``config = hf_model.config.to_dict()`` and ``ckpt = hf_model.state_dict()`` (build_encoder.py:38-45, :71;
build_decoder.py:38-45, :71).  This module reads the same two things straight from the on-disk format —
``config.json`` + ``model.safetensors`` (or sharded ``model-0000x-of-0000y.safetensors`` with its index, or
``pytorch_model.bin``) — without instantiating a torch model, and hands them to ``run.build_encoder`` /
``run.build_decoder`` / ``WhisperEngine``.  Host-side plumbing only; no arithmetic.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Tuple

import torch

# keys of WhisperConfig.to_dict() the path reads (run.py:150-169,273; build_*.py ctor calls) and their defaults
# (configuration_whisper.py:196-235)
_CONFIG_DEFAULTS = dict(
    vocab_size=51865, num_mel_bins=80, d_model=256, encoder_layers=6, decoder_layers=6, encoder_attention_heads=4,
    decoder_attention_heads=4, encoder_ffn_dim=1536, decoder_ffn_dim=1536, max_source_positions=1500,
    max_target_positions=448, activation_function="gelu", scale_embedding=False, pad_token_id=50256, bos_token_id=50257,
    eos_token_id=50256, decoder_start_token_id=50257, suppress_tokens=None, begin_suppress_tokens=[220, 50256],
    forced_decoder_ids=None, forced_bos_token_id=None, max_length=448)

_REQUIRED_PREFIXES = ("model.encoder.conv1.weight", "model.encoder.embed_positions.weight", "model.decoder.embed_tokens.weight",
                      "model.decoder.embed_positions.weight", "model.decoder.layer_norm.weight")


def load_config(path: str) -> Dict:
    """``config.json`` (+ ``generation_config.json`` when present: newer checkpoints keep suppress/forced ids there)."""
    with open(os.path.join(path, "config.json")) as f:
        raw = json.load(f)
    cfg = dict(_CONFIG_DEFAULTS)
    cfg.update({k: v for k, v in raw.items() if v is not None or k in ("forced_bos_token_id",)})
    gen_path = os.path.join(path, "generation_config.json")
    if os.path.exists(gen_path):
        with open(gen_path) as f:
            gen = json.load(f)
        for k in ("suppress_tokens", "begin_suppress_tokens", "forced_decoder_ids", "max_length", "pad_token_id", "eos_token_id",
                  "decoder_start_token_id"):
            if gen.get(k) is not None and (raw.get(k) is None):
                cfg[k] = gen[k]
    if cfg["suppress_tokens"] is None:
        cfg["suppress_tokens"] = []
    if cfg["forced_decoder_ids"] is None:
        raise ValueError("forced_decoder_ids is missing: the greedy session indexes forced_decoder_ids[-1][0] "
                         "(generation/utils.py:896, run.py:157) — for `.en` checkpoints it is [[1, 50362]]")
    if isinstance(cfg["eos_token_id"], list):
        cfg["eos_token_id"] = cfg["eos_token_id"][0]
    if cfg["activation_function"] != "gelu":
        raise ValueError("only the erf GELU of the Whisper checkpoints is on the path")
    return cfg


def _load_safetensors(files) -> Dict[str, torch.Tensor]:
    from safetensors import safe_open
    sd = {}
    for fn in files:
        with safe_open(fn, framework="pt", device="cpu") as f:
            for k in f.keys():
                sd[k] = f.get_tensor(k)
    return sd


def load_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """All tensors of the checkpoint as fp32 CPU tensors with the oracle's key names."""
    single = os.path.join(path, "model.safetensors")
    index = os.path.join(path, "model.safetensors.index.json")
    binf = os.path.join(path, "pytorch_model.bin")
    if os.path.exists(single):
        sd = _load_safetensors([single])
    elif os.path.exists(index):
        with open(index) as f:
            shards = sorted(set(json.load(f)["weight_map"].values()))
        sd = _load_safetensors([os.path.join(path, s) for s in shards])
    elif os.path.exists(binf):
        sd = torch.load(binf, map_location="cpu", weights_only=True)
    else:
        raise FileNotFoundError(f"no model.safetensors / model.safetensors.index.json / pytorch_model.bin under {path}")
    sd = {k: v.to(torch.float32) for k, v in sd.items() if isinstance(v, torch.Tensor)}
    # safetensors drops tied duplicates: proj_out.weight shares storage with embed_tokens (modeling_whisper.py:1335)
    if "proj_out.weight" not in sd and "model.decoder.embed_tokens.weight" in sd:
        sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    missing = [k for k in _REQUIRED_PREFIXES if k not in sd]
    if missing:
        raise KeyError(f"checkpoint is not a Whisper seq2seq model, missing {missing}")
    return sd


def load_hf_checkpoint(path: str) -> Tuple[Dict, Dict[str, torch.Tensor]]:
    """``(config, ckpt)`` exactly as build_encoder.py:38-45 / build_decoder.py:38-45 obtain them."""
    cfg = load_config(path)
    sd = load_state_dict(path)
    d = sd["model.decoder.embed_tokens.weight"]
    if tuple(d.shape) != (cfg["vocab_size"], cfg["d_model"]):
        raise ValueError(f"embed_tokens {tuple(d.shape)} does not match config (vocab {cfg['vocab_size']}, d_model {cfg['d_model']})")
    return cfg, sd


def save_hf_checkpoint(path: str, config: Dict, state_dict: Dict[str, torch.Tensor], safetensors: bool = True):
    """Write a checkpoint directory in the HF layout (used by the tests and to export synthetic models)."""
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump({k: v for k, v in config.items() if k != "name"}, f)
    sd = {k: v.detach().to("cpu").contiguous() for k, v in state_dict.items()}
    if safetensors:
        from safetensors.torch import save_file
        sd.pop("proj_out.weight", None)       # tied: stored once, like transformers' save_pretrained
        save_file(sd, os.path.join(path, "model.safetensors"))
    else:
        torch.save(sd, os.path.join(path, "pytorch_model.bin"))
