"""HF checkpoint loader: config.json + model.safetensors / pytorch_model.bin -> the (config, ckpt)
pair the reference's build scripts get from from_pretrained (build_encoder.py:38-45, :71)."""
import json
import os

import pytest
import torch

from oracle import synth
from whisper_trtllm_b200 import checkpoint, run, runtime


@pytest.fixture(scope="module")
def micro():
    cfg = synth.make_config("micro")
    return cfg, synth.make_weights(cfg, seed=2)


@pytest.mark.parametrize("fmt", ["safetensors", "bin"])
def test_roundtrip_and_tied_head(micro, tmp_path, fmt):
    cfg, sd = micro
    path = str(tmp_path / fmt)
    checkpoint.save_hf_checkpoint(path, cfg, sd, safetensors=(fmt == "safetensors"))
    cfg2, sd2 = checkpoint.load_hf_checkpoint(path)
    for k in ("d_model", "encoder_layers", "decoder_attention_heads", "vocab_size", "suppress_tokens", "forced_decoder_ids",
              "begin_suppress_tokens", "max_length", "decoder_start_token_id"):
        assert cfg2[k] == cfg[k], k
    assert set(sd2) == set(sd)
    for k in sd:
        assert torch.equal(sd2[k], sd[k]), k
    assert torch.equal(sd2["proj_out.weight"], sd2["model.decoder.embed_tokens.weight"])
    # the loaded pair feeds the engine builders unchanged
    buf = run.build_decoder(cfg2, sd2)
    kind, dec = runtime.deserialize_engine(buf)
    assert kind == "WhisperDecoder" and torch.equal(dec.layers[1].fc1.weight.data.cpu(), sd["model.decoder.layers.1.fc1.weight"])


def test_sharded_safetensors_and_generation_config(micro, tmp_path):
    from safetensors.torch import save_file
    cfg, sd = micro
    path = str(tmp_path)
    keys = sorted(k for k in sd if k != "proj_out.weight")
    half = len(keys) // 2
    shards = {"model-00001-of-00002.safetensors": keys[:half], "model-00002-of-00002.safetensors": keys[half:]}
    weight_map = {}
    for fn, ks in shards.items():
        save_file({k: sd[k].contiguous() for k in ks}, os.path.join(path, fn))
        weight_map.update({k: fn for k in ks})
    with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
        json.dump({"weight_map": weight_map}, f)
    # newer checkpoints: suppress / forced ids live in generation_config.json, config.json has nulls
    c = {k: v for k, v in cfg.items() if k not in ("name", "suppress_tokens", "forced_decoder_ids")}
    c["suppress_tokens"] = None
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(c, f)
    with open(os.path.join(path, "generation_config.json"), "w") as f:
        json.dump({"suppress_tokens": cfg["suppress_tokens"], "forced_decoder_ids": cfg["forced_decoder_ids"]}, f)
    cfg2, sd2 = checkpoint.load_hf_checkpoint(path)
    assert cfg2["suppress_tokens"] == cfg["suppress_tokens"] and cfg2["forced_decoder_ids"] == cfg["forced_decoder_ids"]
    assert all(torch.equal(sd2[k], sd[k]) for k in sd)


def test_errors_are_loud(micro, tmp_path):
    cfg, sd = micro
    with pytest.raises(FileNotFoundError):
        checkpoint.load_state_dict(str(tmp_path))
    c = {k: v for k, v in cfg.items() if k not in ("name", "forced_decoder_ids")}
    with open(os.path.join(str(tmp_path), "config.json"), "w") as f:
        json.dump(c, f)
    with pytest.raises(ValueError):
        checkpoint.load_config(str(tmp_path))     # forced_decoder_ids missing (generation/utils.py:896)
    checkpoint.save_hf_checkpoint(str(tmp_path / "bad"), cfg, {"foo": torch.zeros(2)})
    with pytest.raises(KeyError):
        checkpoint.load_state_dict(str(tmp_path / "bad"))


def test_directory_written_by_installed_transformers(tmp_path):
    """The on-disk format as today's `transformers` writes it (`save_pretrained`: tied head stored once, generation fields in
    generation_config.json): every tensor of the model's state_dict comes back under the same key, bit for bit."""
    transformers = pytest.importorskip("transformers")
    cfg = transformers.WhisperConfig(
        vocab_size=51864, num_mel_bins=80, d_model=64, encoder_layers=2, decoder_layers=2, encoder_attention_heads=1,
        decoder_attention_heads=1, encoder_ffn_dim=256, decoder_ffn_dim=256, max_source_positions=1500, max_target_positions=448,
        pad_token_id=50256, bos_token_id=50256, eos_token_id=50256, decoder_start_token_id=50257)
    torch.manual_seed(0)
    model = transformers.WhisperForConditionalGeneration(cfg)
    model.generation_config.forced_decoder_ids = [[1, 50362]]
    model.generation_config.suppress_tokens = [1, 2, 7]
    model.generation_config.max_length = 448
    model.save_pretrained(str(tmp_path))
    cfg2, sd2 = checkpoint.load_hf_checkpoint(str(tmp_path))
    assert (cfg2["d_model"], cfg2["encoder_layers"], cfg2["decoder_attention_heads"], cfg2["vocab_size"]) == (64, 2, 1, 51864)
    assert cfg2["forced_decoder_ids"] == [[1, 50362]] and cfg2["suppress_tokens"] == [1, 2, 7] and cfg2["max_length"] == 448
    assert cfg2["begin_suppress_tokens"] == [220, 50256] and cfg2["decoder_start_token_id"] == 50257
    sd = model.state_dict()
    assert set(sd2) == set(sd)
    for k, v in sd.items():
        assert torch.equal(sd2[k], v.float()), k
    assert torch.equal(sd2["proj_out.weight"], sd2["model.decoder.embed_tokens.weight"])
