"""The CPU oracle restatement (oracle/whisper_ref.py) against vectors generated from the REAL reference
(oracle/make_golden.py ran the vendored simplified-HF Whisper in the build container)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import synth, whisper_ref as R
from oracle.make_golden import CASES, subsample_enc, subsample_logits


@pytest.mark.parametrize("case", ["micro", "tiny_b1"])
def test_restatement_matches_reference_vectors(case):
    meta, g = load_golden(case)
    size, B, wseed, mseed, max_length, _ = CASES[case]
    cfg = synth.make_config(size, max_length=max_length)
    sd = synth.make_weights(cfg, seed=wseed)
    assert synth.weights_fingerprint(sd) == meta["weights_fingerprint"], "synthetic weights are not bit-reproducible here"
    mel = synth.make_mel(B, seed=mseed)
    steps = 40  # bounded so the CPU suite stays fast; the full length is covered by make_golden itself
    ids, enc, logits = R.greedy(mel, sd, cfg, max_new_tokens=steps, return_logits=True)
    assert np.array_equal(ids.numpy(), g["tokens"][:, :steps + 1])
    np.testing.assert_allclose(subsample_enc(enc).numpy(), g["enc_sub"], rtol=1e-4, atol=1e-4)
    for i, s in enumerate(g["logit_steps"]):
        if s < steps:
            np.testing.assert_allclose(subsample_logits(logits[s]).numpy(), g["logits_sub"][i], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("case", sorted(CASES))
def test_golden_metadata(case):
    meta, g = load_golden(case)
    assert meta["restatement_tokens_equal"] and meta["restatement_logit_maxabs"] < 1e-4 and meta["restatement_enc_maxabs"] < 1e-4
    assert g["tokens"].shape[0] == meta["batch"] and g["tokens"].shape[1] <= meta["max_length"]
    assert (g["tokens"][:, 0] == 50257).all() and (g["tokens"][:, 1] == 50362).all()  # sot, forced <|notimestamps|>
    assert min(meta["distinct_tokens_per_row"]) >= 3, "degenerate synthetic model"


def test_logits_processors_order_and_indices():
    cfg = synth.make_config("micro")
    V = cfg["vocab_size"]
    x = torch.zeros(2, V)
    x[:, 220] = 5.0
    x[:, 1] = 9.0     # in suppress list
    x[:, 300] = 4.0
    s = R.process_logits(x, 1, cfg)   # forced token at generation index 1
    assert torch.argmax(s, -1).tolist() == [50362, 50362] and s[0, 50362] == 0 and torch.isinf(s[0, 300])
    s = R.process_logits(x, 2, cfg)   # begin-suppress (220, eos) at begin_index = 2
    assert torch.argmax(s, -1).tolist() == [300, 300]
    s = R.process_logits(x, 3, cfg)
    assert torch.argmax(s, -1).tolist() == [220, 220] and torch.isinf(s[0, 1])


def test_four_attention_modes_consistent():
    """self/cross x with/without cache give the same numbers as recomputing from scratch (MW:474-503)."""
    cfg = synth.make_config("micro")
    sd = synth.make_weights(cfg, seed=3)
    torch.manual_seed(0)
    H = cfg["decoder_attention_heads"]
    enc = torch.randn(2, 1500, cfg["d_model"])
    p = "model.decoder.layers.0"
    h1, h2 = torch.randn(2, 1, cfg["d_model"]), torch.randn(2, 1, cfg["d_model"])
    o1, (k1, v1) = R.decoder_attention(h1, sd, p + ".self_attn", H)                       # self, no cache
    o2, (k2, v2) = R.decoder_attention(h2, sd, p + ".self_attn", H, past=(k1, v1))        # self + cache
    assert k2.shape[2] == 2 and torch.equal(k2[:, :, :1], k1)
    c1, (ck, cv) = R.decoder_attention(h1, sd, p + ".encoder_attn", H, enc)               # cross, no cache
    c2, (ck2, cv2) = R.decoder_attention(h1, sd, p + ".encoder_attn", H, enc, past=(ck, cv))  # cross + cache
    assert ck2 is ck and torch.allclose(c1, c2, atol=0, rtol=0)
    assert ck.shape == (2, H, 1500, 64)
