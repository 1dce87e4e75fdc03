"""`whisper_trtllm_b200.processor.WhisperProcessor`: the two calls the reference's scripts make on its processor (run.py:239, :267,
:287) — argument conventions and errors of WhisperFeatureExtractor.__call__ (feature_extraction_whisper.py:136-260), decode through
the detokenizer.  The GPU front-end is replaced by a stub here; its arithmetic is covered by tests/test_logmel.py (-m gpu)."""
import json

import numpy as np
import pytest
import torch

from whisper_trtllm_b200 import processor as P
from whisper_trtllm_b200.text import bytes_to_unicode


class StubFrontend:
    def __init__(self):
        self.calls = []

    def __call__(self, waves):
        self.calls.append([np.asarray(w) for w in waves])
        return torch.zeros(len(waves), 80, 3000)


def test_input_conventions_follow_the_reference():
    one = np.arange(5, dtype=np.float64)
    assert [a.tolist() for a in P.as_batch(one)] == [[0, 1, 2, 3, 4]] and P.as_batch(one)[0].dtype == np.float32
    assert len(P.as_batch([0.1, 0.2, 0.3])) == 1                                   # a list of floats is ONE waveform
    assert [len(a) for a in P.as_batch([np.zeros(3), np.zeros(7)])] == [3, 7]      # list of arrays: ragged batch
    assert [len(a) for a in P.as_batch([[0.0, 1.0], [2.0, 3.0, 4.0]])] == [2, 3]   # list of lists
    assert [len(a) for a in P.as_batch(np.zeros((4, 9)))] == [9] * 4               # 2-D array: batch
    assert [len(a) for a in P.as_batch(torch.zeros(2, 6))] == [6, 6]
    with pytest.raises(ValueError, match="mono"):
        P.as_batch(np.zeros((2, 2, 9)))
    with pytest.raises(ValueError, match="mono"):
        P.as_batch([np.zeros((2, 9))])


def test_feature_extractor_call_and_errors():
    stub = StubFrontend()
    fe = P.WhisperFeatureExtractor(frontend=stub)
    out = fe(np.zeros(1000), sampling_rate=16000, return_tensors="pt")
    assert out.input_features.shape == (1, 80, 3000) and out["input_features"] is out.input_features
    assert len(stub.calls) == 1 and stub.calls[0][0].shape == (1000,)
    assert fe([np.zeros(10), np.zeros(20)], sampling_rate=16000).input_features.shape[0] == 2
    with pytest.raises(ValueError, match="sampling rate of 16000"):
        fe(np.zeros(10), sampling_rate=8000)                                       # feature_extraction_whisper.py:195-201
    for kw in (dict(do_normalize=True), dict(padding="longest"), dict(truncation=False), dict(max_length=1000),
               dict(return_attention_mask=True), dict(return_tensors="np"), dict(pad_to_multiple_of=8), dict(foo=1)):
        with pytest.raises(NotImplementedError):
            fe(np.zeros(10), sampling_rate=16000, **kw)
    with pytest.raises(AttributeError):
        out.attention_mask
    assert (fe.sampling_rate, fe.n_samples, fe.feature_size, fe.hop_length, fe.n_fft) == (16000, 480000, 80, 160, 400)


def test_processor_from_a_checkpoint_directory(tmp_path):
    from oracle import synth
    from whisper_trtllm_b200 import checkpoint
    cfg = synth.make_config("micro")
    checkpoint.save_hf_checkpoint(str(tmp_path), cfg, synth.make_weights(cfg, seed=2))
    proc = P.WhisperProcessor.from_pretrained(str(tmp_path))
    with pytest.raises(FileNotFoundError):
        proc.batch_decode([[1, 2]])                                                # no vocab.json yet
    b2u = bytes_to_unicode()
    enc = lambda s: "".join(b2u[b] for b in s.encode())
    vocab = {enc(" hi"): 0, enc(" there"): 1, enc(" ."): 2, "<|endoftext|>": 50256}
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    (tmp_path / "added_tokens.json").write_text(json.dumps({"<|startoftranscript|>": 50257, "<|notimestamps|>": 50362}))
    proc = P.WhisperProcessor.from_pretrained(str(tmp_path))
    ids = torch.tensor([[50257, 50362, 0, 1, 2, 50256, 50256]])
    assert proc.batch_decode(ids, skip_special_tokens=True) == [" hi there."]      # fast-tokenizer clean-up: " ." -> "."
    assert proc.decode(ids[0], skip_special_tokens=True, clean_up_tokenization_spaces=False) == " hi there ."
    assert proc.batch_decode(ids)[0].startswith("<|startoftranscript|><|notimestamps|> hi")   # default keeps the special tokens
    (tmp_path / "tokenizer_config.json").write_text(json.dumps({"clean_up_tokenization_spaces": False}))
    assert P.WhisperProcessor.from_pretrained(str(tmp_path)).batch_decode(ids, skip_special_tokens=True) == [" hi there ."]
    proc.feature_extractor._frontend = StubFrontend()
    assert proc(np.zeros(100), sampling_rate=16000, return_tensors="pt").input_features.shape == (1, 80, 3000)
    with pytest.raises(ValueError):
        proc()


def test_no_cpu_feature_path():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from whisper_trtllm_b200 import WhisperB200Error
    with pytest.raises(WhisperB200Error):
        P.WhisperFeatureExtractor()(np.zeros(100), sampling_rate=16000)
