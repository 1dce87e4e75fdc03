"""CPU checks of the host-side mirror of the reference interface (no kernels run here): constructor signatures and
parameter names the reference's binders use, the binder mapping and its inverse, the engine container, the Session
shape contract, the logits processors / stopping criteria, and the generic greedy loop's bookkeeping."""
import inspect

import numpy as np
import pytest
import torch

from oracle import synth, whisper_ref as R
from whisper_trtllm_b200 import models, run, runtime
from whisper_trtllm_b200.model import (decoder_from_config, encoder_from_config, export_hf_state_dict, load_decoder_from_hf,
                                       load_encoder_from_hf)


@pytest.fixture(scope="module")
def micro():
    cfg = synth.make_config("micro")
    sd = synth.make_weights(cfg, seed=0)
    return cfg, sd


def test_constructor_signatures_match_reference():
    # tensorrt_llm/models/whisper/model.py:69-72, 372-382, 155-170, 307
    enc = inspect.signature(models.WhisperEncoder.__init__).parameters
    assert list(enc)[1:8] == ["d_model", "num_mel_bins", "max_source_positions", "encoder_layers", "encoder_attention_heads",
                              "activation_function", "encoder_ffn_dim"]
    dec = inspect.signature(models.WhisperDecoder.__init__).parameters
    assert list(dec)[1:11] == ["pad_token_id", "max_target_positions", "max_source_positions", "d_model", "scale_embedding",
                               "vocab_size", "decoder_layers", "decoder_attention_heads", "activation_function", "decoder_ffn_dim"]
    att = inspect.signature(models.WhisperDecoderAttention.forward).parameters
    assert list(att)[1:6] == ["hidden_states", "key_value_states", "past_key", "past_value", "cache_mask"]
    lay = inspect.signature(models.WhisperDecoderLayer.forward).parameters
    assert list(lay)[1:] == ["hidden_states", "encoder_hidden_states", "self_past_key", "self_past_value", "self_cache_mask",
                             "cross_past_key", "cross_past_value", "cross_cache_mask"]
    fwd = inspect.signature(models.WhisperDecoder.forward).parameters
    assert list(fwd)[1:] == ["input_ids", "encoder_hidden_states", "past_self_keys", "past_self_values", "past_cross_keys",
                             "past_cross_values", "past_self_cache_mask", "past_cross_cache_mask"]
    gs = inspect.signature(run.greedy_search).parameters
    assert list(gs) == ["model", "encoder_outputs", "input_ids", "logits_processor", "stopping_criteria", "pad_token_id", "eos_token_id"]


def test_binder_attribute_names_resolve(micro):
    """Every attribute the reference's binders assign (build_encoder.py:72-91, build_decoder.py:72-101) exists."""
    cfg, sd = micro
    enc, dec = encoder_from_config(cfg), decoder_from_config(cfg)
    enc.conv1.weight.value = sd["model.encoder.conv1.weight"].unsqueeze(2).numpy()
    assert tuple(enc.conv1.weight.shape) == (cfg["d_model"], 80, 1, 3)
    enc.embed_positions_weight = sd["model.encoder.embed_positions.weight"].unsqueeze(0).numpy()
    l0 = enc.layers[0]
    for name in ("qkv", "dense"):
        assert hasattr(getattr(l0.self_attn, name), "weight") and hasattr(getattr(l0.self_attn, name), "bias")
    for name in ("self_attn_layer_norm", "final_layer_norm", "fc1", "fc2"):
        assert hasattr(getattr(l0, name), "weight")
    d0 = dec.layers[0]
    for attn in ("self_attn", "encoder_attn"):
        a = getattr(d0, attn)
        assert a.k_proj.bias is None and a.q_proj.bias is not None and a.v_proj.bias is not None and a.dense.bias is not None
    for name in ("embed_tokens", "embed_positions", "layer_norm", "proj_out"):
        assert hasattr(getattr(dec, name), "weight")
    assert dec.proj_out.bias is None
    names = dict(dec.named_parameters())
    assert "layers.1.encoder_attn.dense.weight" in names and "layers.0.encoder_attn_layer_norm.bias" in names


def test_binder_roundtrip_and_zero_k_bias(micro):
    cfg, sd = micro
    enc = load_encoder_from_hf(encoder_from_config(cfg), sd)
    dec = load_decoder_from_hf(decoder_from_config(cfg), sd)
    d = cfg["d_model"]
    b = enc.layers[0].self_attn.qkv.bias.data
    assert torch.count_nonzero(b[d:2 * d]) == 0 and torch.count_nonzero(b[:d]) > 0   # build_encoder.py:79
    back = export_hf_state_dict(enc, dec)
    assert set(back) == set(sd)
    for k in sd:
        assert torch.equal(back[k].cpu().reshape(sd[k].shape), sd[k]), k


def test_engine_container_roundtrip(micro, tmp_path):
    cfg, sd = micro
    buf = run.build_encoder(cfg, sd, str(tmp_path), "float32")
    assert buf[:8] == runtime.MAGIC and (tmp_path / "WhisperEncoder.engine").exists() and (tmp_path / "config.pkl").exists()
    kind, enc = runtime.deserialize_engine(buf)
    assert kind == "WhisperEncoder" and len(enc.layers) == cfg["encoder_layers"]
    assert np.array_equal(enc.embed_positions_weight[0], sd["model.encoder.embed_positions.weight"].numpy())
    assert torch.equal(enc.layers[1].fc2.weight.data.cpu(), sd["model.encoder.layers.1.fc2.weight"])
    dbuf = run.build_decoder(cfg, sd, str(tmp_path), "bfloat16")
    kind, dec = runtime.deserialize_engine(dbuf)
    assert kind == "WhisperDecoder" and dec.dtype == torch.bfloat16
    assert torch.equal(dec.proj_out.weight.data.cpu(), sd["model.decoder.embed_tokens.weight"])
    with pytest.raises(Exception):
        runtime.deserialize_engine(b"not an engine, e.g. a TensorRT plan")
    with pytest.raises(Exception):
        runtime.deserialize_engine(buf[:len(buf) // 2])


def test_session_shape_contract(micro):
    """infer_shapes follows model.py:464-516: cache lengths come from the MASK SHAPES."""
    cfg, sd = micro
    L, H, d, V = cfg["decoder_layers"], cfg["decoder_attention_heads"], cfg["d_model"], cfg["vocab_size"]
    T = runtime.TensorInfo
    f32, i32 = runtime.DataType.float32, runtime.DataType.int32
    enc_s = runtime.Session.from_serialized_engine(run.build_encoder(cfg, sd))
    out = enc_s.infer_shapes([T("data", f32, (3, 80, 3000)), T("length", f32, (3,))])
    assert [(o.name, tuple(o.shape)) for o in out] == [("hidden_states", (3, 1500, d))]
    assert enc_s.infer_shapes([T("bogus", f32, (1,))]) is None
    assert enc_s.infer_shapes([T("data", i32, (1, 80, 3000))]) is None          # wrong dtype (session.py:135-137)
    dec_s = runtime.Session.from_serialized_engine(run.build_decoder(cfg, sd))

    def shapes(past_t, self_mask, cross_mask, lead=(L, H), B=1):
        ins = [T("data", i32, (B, 1)), T("length", i32, (B,)), T("encoder_hidden_states", f32, (B, 1500, d)),
               T("self_past_key", f32, lead + (past_t, 64)), T("self_past_value", f32, lead + (past_t, 64)),
               T("cross_past_key", f32, lead + (1500, 64)), T("cross_past_value", f32, lead + (1500, 64)),
               T("past_self_cache_mask", f32, (self_mask,)), T("past_cross_cache_mask", f32, (cross_mask,))]
        return {o.name: tuple(o.shape) for o in dec_s.infer_shapes(ins)}
    s0 = shapes(1, 1, 1)                       # step 0 of run.py:108-119: dummy T=1 caches, masks of length 1
    assert s0["hidden_states"] == (1, 1, V) and s0["next_self_keys"] == (L, H, 1, 64) and s0["next_cross_values"] == (L, H, 1500, 64)
    s5 = shapes(5, 6, 1501)
    assert s5["next_self_keys"] == (L, H, 6, 64)
    sb = shapes(7, 8, 1501, lead=(L, 4, H), B=4)   # batched layout [L, B, H, T, 64]
    assert sb["next_self_values"] == (L, 4, H, 8, 64) and sb["hidden_states"] == (4, 1, V)


def test_logits_processors_match_oracle(micro):
    cfg, _ = micro
    procs = run.get_logits_processor(cfg, 1)
    assert procs[1].begin_index == 2                                   # run.py:155-158 for .en models
    g = torch.Generator().manual_seed(0)
    for n in (1, 2, 3, 17):
        scores = torch.randn(3, cfg["vocab_size"], generator=g)
        ids = torch.zeros(3, n, dtype=torch.long)
        want = R.process_logits(scores, n, cfg)
        got = procs(ids, scores.clone())
        assert torch.equal(got, want), n
    crit = run.get_stopping_criteria(cfg)
    assert crit.max_length == 448 and not crit(torch.zeros(1, 447), None) and crit(torch.zeros(1, 448), None)


def test_generic_greedy_loop_bookkeeping(micro, monkeypatch):
    """run.greedy_search's generic path (host logic: pad after EOS, early stop, max_length) against the oracle's loop,
    using the CPU oracle as the ``model`` callable and torch.argmax in place of the CUDA argmax kernel."""
    cfg, sd = micro
    monkeypatch.setattr(run, "_argmax", lambda s: torch.argmax(s, dim=-1).to(torch.int32))
    mel = synth.make_mel(3, seed=1234)
    free = R.greedy(mel, sd, cfg, max_new_tokens=12)
    eos = int(free[0, 4])                      # make row 0 finish early
    cfg2 = dict(cfg, eos_token_id=eos, pad_token_id=eos, max_length=13)
    want = R.greedy(mel, sd, cfg2)
    enc = R.encode(mel, sd, cfg2)

    def model(last_ids, encoder_outputs, past):
        logits, new_past = R.decoder_forward(last_ids.long(), encoder_outputs, sd, cfg2, past)
        return logits, new_past
    ids0 = torch.full((3, 1), cfg2["decoder_start_token_id"], dtype=torch.int32)
    got = run.greedy_search(model, enc, ids0, run.get_logits_processor(cfg2, 1), run.get_stopping_criteria(cfg2),
                            cfg2["pad_token_id"], cfg2["eos_token_id"])
    assert torch.equal(got.long(), want)
    assert (want[0, 5:] == eos).all()          # padded after EOS


def test_forward_without_cuda_fails_loudly(micro):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    cfg, sd = micro
    enc = load_encoder_from_hf(encoder_from_config(cfg), sd)
    from whisper_trtllm_b200 import WhisperB200Error
    with pytest.raises(WhisperB200Error):
        enc(torch.zeros(1, 80, 3000))
