"""fp32 CPU restatement of the reference's Whisper greedy path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Citations: ``MW`` = /root/reference/transformers/src/transformers/models/whisper/modeling_whisper.py,
``GU`` = /root/reference/transformers/src/transformers/generation/utils.py,
``LP`` = /root/reference/transformers/src/transformers/generation/logits_process.py,
``SC`` = /root/reference/transformers/src/transformers/generation/stopping_criteria.py.

The arithmetic of the real oracle is PyTorch ATen CPU kernels (torch 2.11.0 here; not vendored in the
reference tree) — this file calls the same ``torch.nn.functional`` ops on the same
``state_dict`` tensors, so encoder output / logits / token ids can be compared 1:1.  Validated against
the real reference by oracle/make_golden.py (max |diff| recorded in tests/golden/*.json).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def layer_norm(x: torch.Tensor, sd: SD, prefix: str) -> torch.Tensor:
    # nn.LayerNorm(d), eps 1e-5 (MW:608,613 / :675-686)
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def _heads(t: torch.Tensor, n_heads: int) -> torch.Tensor:
    # _shape(): view(B, S, H, 64).transpose(1, 2) (MW:451-452)
    b, s, d = t.shape
    return t.view(b, s, n_heads, d // n_heads).transpose(1, 2).contiguous()


def _attend(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """q [B,H,Tq,64] (already scaled), k/v [B,H,Tk,64] -> [B,Tq,d].  bmm / softmax / bmm, no mask
    (MW:507-523 decoder, MW:576-591 encoder)."""
    b, h, tq, dh = q.shape
    w = torch.bmm(q.reshape(b * h, tq, dh), k.reshape(b * h, -1, dh).transpose(1, 2))
    w = F.softmax(w, dim=-1)
    o = torch.bmm(w, v.reshape(b * h, -1, dh))
    return o.view(b, h, tq, dh).transpose(1, 2).reshape(b, tq, h * dh)


def encoder_stem(mel: torch.Tensor, sd: SD) -> torch.Tensor:
    """MW:992-997 — gelu(conv1) -> gelu(conv2, stride 2) -> permute -> + embed_positions.weight."""
    p = "model.encoder."
    h = F.gelu(F.conv1d(mel, sd[p + "conv1.weight"], sd[p + "conv1.bias"], stride=1, padding=1))
    h = F.gelu(F.conv1d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], stride=2, padding=1))
    return h.permute(0, 2, 1) + sd[p + "embed_positions.weight"]


def encoder_attention(x: torch.Tensor, sd: SD, prefix: str, n_heads: int) -> torch.Tensor:
    """WhisperEncoderAttention.forward, MW:569-593.  q scaled BEFORE QK^T (:572); k_proj has no bias (:547)."""
    scaling = (x.shape[-1] // n_heads) ** -0.5
    q = F.linear(x, sd[prefix + ".q_proj.weight"], sd[prefix + ".q_proj.bias"]) * scaling
    k = F.linear(x, sd[prefix + ".k_proj.weight"])
    v = F.linear(x, sd[prefix + ".v_proj.weight"], sd[prefix + ".v_proj.bias"])
    o = _attend(_heads(q, n_heads), _heads(k, n_heads), _heads(v, n_heads))
    return F.linear(o, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def encoder_layer(x: torch.Tensor, sd: SD, prefix: str, n_heads: int) -> torch.Tensor:
    """WhisperEncoderLayer.forward, MW:632-643 (pre-LN attn + residual, pre-LN erf-GELU MLP + residual)."""
    x = x + encoder_attention(layer_norm(x, sd, prefix + ".self_attn_layer_norm"), sd, prefix + ".self_attn", n_heads)
    m = layer_norm(x, sd, prefix + ".final_layer_norm")
    m = F.gelu(F.linear(m, sd[prefix + ".fc1.weight"], sd[prefix + ".fc1.bias"]))  # activations.py:214 erf GELU
    m = F.linear(m, sd[prefix + ".fc2.weight"], sd[prefix + ".fc2.bias"])
    return x + m


def encode(mel: torch.Tensor, sd: SD, cfg: Dict) -> torch.Tensor:
    """WhisperEncoder.forward, MW:992-1011: mel [B,80,3000] -> [B,1500,d]."""
    x = encoder_stem(mel, sd)
    for i in range(cfg["encoder_layers"]):
        x = encoder_layer(x, sd, f"model.encoder.layers.{i}", cfg["encoder_attention_heads"])
    return layer_norm(x, sd, "model.encoder.layer_norm")


def cross_kv(enc: torch.Tensor, sd: SD, cfg: Dict) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Cross-attention K/V, computed once (branch MW:484-487): K = enc Wk^T (no bias), V = enc Wv^T + bv,
    each [B,H,1500,64] per decoder layer."""
    H = cfg["decoder_attention_heads"]
    out = []
    for i in range(cfg["decoder_layers"]):
        p = f"model.decoder.layers.{i}.encoder_attn"
        k = _heads(F.linear(enc, sd[p + ".k_proj.weight"]), H)
        v = _heads(F.linear(enc, sd[p + ".v_proj.weight"], sd[p + ".v_proj.bias"]), H)
        out.append((k, v))
    return out


def decoder_attention(
    hidden: torch.Tensor,
    sd: SD,
    prefix: str,
    n_heads: int,
    key_value_states: Optional[torch.Tensor] = None,
    past: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """WhisperDecoderAttention.forward with its FOUR branches, MW:468-526.

    * cross + cache (``past[0].shape[2] == key_value_states.shape[1]``): reuse past      (:474-481)
    * cross, no cache: project key_value_states                                         (:484-487)
    * self + cache: cat(past, proj(hidden)) on dim 2                                    (:490-495)
    * self, no cache: proj(hidden)                                                      (:498-501)
    No attention mask of any kind (tgt_len is 1 in the greedy loop); q pre-scaled (:472).
    """
    scaling = (hidden.shape[-1] // n_heads) ** -0.5
    q = F.linear(hidden, sd[prefix + ".q_proj.weight"], sd[prefix + ".q_proj.bias"]) * scaling
    is_cross = key_value_states is not None
    if is_cross and past is not None and past[0].shape[2] == key_value_states.shape[1]:
        k, v = past
    elif is_cross:
        k = _heads(F.linear(key_value_states, sd[prefix + ".k_proj.weight"]), n_heads)
        v = _heads(F.linear(key_value_states, sd[prefix + ".v_proj.weight"], sd[prefix + ".v_proj.bias"]), n_heads)
    else:
        k = _heads(F.linear(hidden, sd[prefix + ".k_proj.weight"]), n_heads)
        v = _heads(F.linear(hidden, sd[prefix + ".v_proj.weight"], sd[prefix + ".v_proj.bias"]), n_heads)
        if past is not None:
            k = torch.cat([past[0], k], dim=2)
            v = torch.cat([past[1], v], dim=2)
    o = _attend(_heads(q, n_heads), k, v)
    return F.linear(o, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"]), (k, v)


def decoder_layer(x, sd: SD, prefix: str, n_heads: int, enc, past):
    """WhisperDecoderLayer.forward, MW:710-751.  past = (self_k, self_v, cross_k, cross_v) or None."""
    a, (sk, sv) = decoder_attention(
        layer_norm(x, sd, prefix + ".self_attn_layer_norm"), sd, prefix + ".self_attn", n_heads,
        None, past[:2] if past is not None else None)
    x = x + a
    a, (ck, cv) = decoder_attention(
        layer_norm(x, sd, prefix + ".encoder_attn_layer_norm"), sd, prefix + ".encoder_attn", n_heads,
        enc, past[-2:] if past is not None else None)
    x = x + a
    m = layer_norm(x, sd, prefix + ".final_layer_norm")
    m = F.gelu(F.linear(m, sd[prefix + ".fc1.weight"], sd[prefix + ".fc1.bias"]))
    m = F.linear(m, sd[prefix + ".fc2.weight"], sd[prefix + ".fc2.bias"])
    return x + m, (sk, sv, ck, cv)


def decoder_forward(ids: torch.Tensor, enc: torch.Tensor, sd: SD, cfg: Dict, past=None):
    """WhisperDecoder.forward (MW:1143-1185) + proj_out (MW:1433): ids [B,T] -> logits [B,T,V], new past.

    Position row = past length (WhisperPositionalEmbedding.forward, MW:307-308); no embed_scale multiply.
    """
    past_len = past[0][0].shape[2] if past is not None else 0
    x = F.embedding(ids, sd["model.decoder.embed_tokens.weight"])
    x = x + sd["model.decoder.embed_positions.weight"][past_len:past_len + ids.shape[1]]
    new_past = []
    for i in range(cfg["decoder_layers"]):
        x, p = decoder_layer(x, sd, f"model.decoder.layers.{i}", cfg["decoder_attention_heads"], enc,
                             past[i] if past is not None else None)
        new_past.append(p)
    x = layer_norm(x, sd, "model.decoder.layer_norm")
    return F.linear(x, sd["proj_out.weight"]), new_past


def process_logits(scores: torch.Tensor, cur_len: int, cfg: Dict) -> torch.Tensor:
    """The three processors in the reference's order (GU:890-899, run.py:150-162):
    SuppressTokens (LP:1300-1310) -> SuppressTokensAtBegin (LP:1281-1297) -> ForceTokens (LP:1313-1328).
    begin_index = 1 (+1 if forced_bos_token_id) + forced_decoder_ids[-1][0]   (GU:894-897, run.py:155-158).
    """
    scores = scores.clone()
    scores[:, cfg["suppress_tokens"]] = -float("inf")
    begin_index = 1
    if cfg.get("forced_bos_token_id") is not None:
        begin_index += 1
    begin_index += cfg["forced_decoder_ids"][-1][0]
    if cur_len == begin_index:
        scores[:, cfg["begin_suppress_tokens"]] = -float("inf")
    tok = dict(cfg["forced_decoder_ids"]).get(cur_len, None)
    if tok is not None:
        scores[:, :] = -float("inf")
        scores[:, tok] = 0
    return scores


@torch.no_grad()
def greedy(mel: torch.Tensor, sd: SD, cfg: Dict, max_new_tokens: Optional[int] = None, return_logits: bool = False):
    """Greedy transcription, GU:1474-1529 (identical to run.py:171-227).

    Returns ids [B, <=max_length] int64 (and the raw per-step logits [steps][B,V] when asked).
    Stop rule: all rows finished (GU:1519-1520) or ``len(ids) >= max_length`` (SC:61-70).
    """
    B = mel.shape[0]
    enc = encode(mel, sd, cfg)
    ids = torch.full((B, 1), cfg["decoder_start_token_id"], dtype=torch.long)
    unfinished = torch.ones(B, dtype=torch.long)
    pad, eos = cfg["pad_token_id"], cfg["eos_token_id"]
    max_length = cfg["max_length"] if max_new_tokens is None else min(cfg["max_length"], 1 + max_new_tokens)
    past = None
    all_logits = []
    while True:
        logits, past = decoder_forward(ids[:, -1:] if past is not None else ids, enc, sd, cfg, past)
        next_logits = logits[:, -1, :]
        if return_logits:
            all_logits.append(next_logits.clone())
        scores = process_logits(next_logits, ids.shape[1], cfg)
        nxt = torch.argmax(scores, dim=-1)
        nxt = nxt * unfinished + pad * (1 - unfinished)
        ids = torch.cat([ids, nxt[:, None]], dim=-1)
        unfinished = unfinished * (nxt != eos).long()
        if unfinished.max() == 0 or ids.shape[1] >= max_length:
            break
    if return_logits:
        return ids, enc, all_logits
    return ids
