"""Seeded synthetic Whisper `.en` configs, weights and log-mel inputs (data generation for bench.py, the tests and the
profiling tools; no arithmetic of the path lives here, and the oracle re-exports it as ``oracle.synth``).

No checkpoint or dataset exists offline, so parity is pinned on synthetic models
whose *state_dict keys and shapes* are exactly the oracle's (dumped from
``WhisperForConditionalGeneration(cfg).state_dict()``).

Bit-determinism across machines: values are built from ``torch.randint`` (integer Mersenne twister)
followed by exact fp32 arithmetic, so the GPU box regenerates bit-identical weights from the seed
(no libm / SIMD normal-sampler differences).  ``weights_fingerprint`` is stored in the goldens.
"""
from __future__ import annotations

import hashlib
from typing import Dict

import torch

# /root/reference/transformers/src/transformers/models/whisper/configuration_whisper.py:37-47
NON_SPEECH_TOKENS = [
    1, 2, 7, 8, 9, 10, 14, 25,
    26, 27, 28, 29, 31, 58, 59, 60, 61, 62,
    63, 90, 91, 92, 93, 357, 366, 438, 532, 685,
    705, 796, 930, 1058, 1220, 1267, 1279, 1303, 1343, 1377,
    1391, 1635, 1782, 1875, 2162, 2361, 2488, 3467, 4008, 4211,
    4600, 4808, 5299, 5855, 6329, 7203, 9609, 9959, 10563, 10786,
    11420, 11709, 11907, 13163, 13697, 13700, 14808, 15306, 16410, 16791,
    17992, 19203, 19510, 20724, 22305, 22935, 27007, 30109, 30420, 33409,
    34949, 40283, 40493, 40549, 47282, 49146, 50257, 50359, 50360, 50361,
]

# size table: SURVEY.md §8 (d_model, layers, heads)
SIZES = {
    "tiny.en": (384, 4, 6),
    "base.en": (512, 6, 8),
    "small.en": (768, 12, 12),
    "medium.en": (1024, 24, 16),
}


def make_config(size: str = "tiny.en", **overrides) -> Dict:
    """Config dict with the keys of the reference's ``config.pkl`` (= ``WhisperConfig.to_dict()``,
    build_encoder.py:42-45) that the path reads (run.py:150-169,273,283-284).

    ``size`` may also be ``"micro"`` (a 2-layer d=128 model for fast CPU tests; not a Whisper size).
    """
    if size == "micro":
        d, L, H = 128, 2, 2
    else:
        d, L, H = SIZES[size]
    cfg = dict(
        name=size,
        vocab_size=51864,
        num_mel_bins=80,
        d_model=d,
        encoder_layers=L,
        decoder_layers=L,
        encoder_attention_heads=H,
        decoder_attention_heads=H,
        encoder_ffn_dim=4 * d,
        decoder_ffn_dim=4 * d,
        max_source_positions=1500,
        max_target_positions=448,
        activation_function="gelu",
        scale_embedding=False,
        pad_token_id=50256,
        bos_token_id=50256,
        eos_token_id=50256,
        decoder_start_token_id=50257,
        suppress_tokens=list(NON_SPEECH_TOKENS),
        begin_suppress_tokens=[220, 50256],
        forced_decoder_ids=[[1, 50362]],
        forced_bos_token_id=None,
        max_length=448,
    )
    cfg.update(overrides)
    return cfg


def _uniform(gen: torch.Generator, shape, std: float) -> torch.Tensor:
    """Zero-mean uniform with the given std, built from 23-bit integers with exact fp32 steps."""
    r = torch.randint(0, 1 << 23, tuple(shape), generator=gen, dtype=torch.int32)
    u = (r.to(torch.float32) + 0.5) * (1.0 / float(1 << 23)) - 0.5  # exact, in (-0.5, 0.5)
    return u * float(std * 12.0 ** 0.5)


def make_weights(cfg: Dict, seed: int = 0, gain: float = 1.5, emb_std: float = 0.1) -> Dict[str, torch.Tensor]:
    """Synthetic state_dict with the oracle's keys.

    Fan-in scaled init (HF's default std=0.02 init gives a degenerate model that emits one token for
    every input, SURVEY.md §7 "Degenerate synthetic models").
    """
    gen = torch.Generator().manual_seed(10_000 + seed)
    d = cfg["d_model"]
    F_enc, F_dec = cfg["encoder_ffn_dim"], cfg["decoder_ffn_dim"]
    V = cfg["vocab_size"]
    sd: Dict[str, torch.Tensor] = {}

    def lin(prefix, n_out, n_in, bias=True, g=gain):
        sd[prefix + ".weight"] = _uniform(gen, (n_out, n_in), g / n_in ** 0.5)
        if bias:
            sd[prefix + ".bias"] = _uniform(gen, (n_out,), 0.02)

    def ln(prefix):
        sd[prefix + ".weight"] = 1.0 + _uniform(gen, (d,), 0.1)
        sd[prefix + ".bias"] = _uniform(gen, (d,), 0.1)

    def attn(prefix):
        lin(prefix + ".k_proj", d, d, bias=False)
        lin(prefix + ".v_proj", d, d)
        lin(prefix + ".q_proj", d, d)
        lin(prefix + ".out_proj", d, d)

    nm = cfg["num_mel_bins"]
    sd["model.encoder.conv1.weight"] = _uniform(gen, (d, nm, 3), gain / (3 * nm) ** 0.5)
    sd["model.encoder.conv1.bias"] = _uniform(gen, (d,), 0.02)
    sd["model.encoder.conv2.weight"] = _uniform(gen, (d, d, 3), gain / (3 * d) ** 0.5)
    sd["model.encoder.conv2.bias"] = _uniform(gen, (d,), 0.02)
    sd["model.encoder.embed_positions.weight"] = _uniform(gen, (cfg["max_source_positions"], d), 0.3)
    for i in range(cfg["encoder_layers"]):
        p = f"model.encoder.layers.{i}"
        attn(p + ".self_attn")
        ln(p + ".self_attn_layer_norm")
        lin(p + ".fc1", F_enc, d)
        lin(p + ".fc2", d, F_enc)
        ln(p + ".final_layer_norm")
    ln("model.encoder.layer_norm")

    sd["model.decoder.embed_tokens.weight"] = _uniform(gen, (V, d), emb_std)
    sd["model.decoder.embed_positions.weight"] = _uniform(gen, (cfg["max_target_positions"], d), emb_std)
    for i in range(cfg["decoder_layers"]):
        p = f"model.decoder.layers.{i}"
        attn(p + ".self_attn")
        ln(p + ".self_attn_layer_norm")
        attn(p + ".encoder_attn")
        ln(p + ".encoder_attn_layer_norm")
        lin(p + ".fc1", F_dec, d)
        lin(p + ".fc2", d, F_dec)
        ln(p + ".final_layer_norm")
    ln("model.decoder.layer_norm")
    # proj_out.weight is the SAME storage as embed_tokens (modeling_whisper.py:1335, modeling_utils.py:1290-1297)
    sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    return sd


def make_mel(batch: int, seed: int = 1234, n_mels: int = 80, frames: int = 3000) -> torch.Tensor:
    """Synthetic log-mel ``U(-1, 1)`` fp32 ``[B, 80, 3000]``, bit-deterministic."""
    gen = torch.Generator().manual_seed(seed)
    r = torch.randint(0, 1 << 23, (batch, n_mels, frames), generator=gen, dtype=torch.int32)
    return (r.to(torch.float32) + 0.5) * (2.0 / float(1 << 23)) - 1.0


def weights_fingerprint(sd: Dict[str, torch.Tensor]) -> str:
    """sha256 over the bytes of a few tensors — proves the GPU box regenerated the same weights."""
    h = hashlib.sha256()
    for k in sorted(sd):
        if k.endswith("layers.0.fc1.weight") or k.endswith("conv1.weight") or k.endswith("layer_norm.bias"):
            h.update(k.encode())
            h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()
