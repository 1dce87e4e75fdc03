"""The whole chain on the GPU (run.py:259-290 / cal_wer.py:251-287 equivalents): .wav files -> GPU log-mel -> encoder + greedy loop
-> ids -> text -> WER, through WhisperPipeline, from a checkpoint directory in the HF layout (synthetic weights, toy vocabulary).
The ids must be those of the CPU oracle run on the same log-mel (fp32: bit-identical)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import logmel_ref as LM
from oracle import synth, whisper_ref as R
from whisper_trtllm_b200 import audio, checkpoint
from whisper_trtllm_b200.text import bytes_to_unicode

pytestmark = pytest.mark.gpu


def _checkpoint_dir(tmp_path, cfg, sd):
    d = str(tmp_path / "ckpt")
    checkpoint.save_hf_checkpoint(d, cfg, sd)
    b2u = bytes_to_unicode()
    pieces = [f" w{i}" for i in range(cfg["eos_token_id"])]                                           # every id decodes
    pieces[:9] = [" zero", " one", " two", " three", " four", " five", " six", " seven", " colour"]   # a few real words
    vocab = {"".join(b2u[b] for b in w.encode()): i for i, w in enumerate(pieces)}
    vocab["<|endoftext|>"] = cfg["eos_token_id"]
    with open(os.path.join(d, "vocab.json"), "w") as f:
        json.dump(vocab, f)
    with open(os.path.join(d, "normalizer.json"), "w") as f:
        json.dump({"colour": "color"}, f)
    return d


def test_wav_files_to_text_and_wer(tmp_path):
    from whisper_trtllm_b200.pipeline import WhisperPipeline
    cfg = synth.make_config("micro", max_length=24)
    sd = synth.make_weights(cfg, seed=4)      # smallest top-1 / top-2 logit gap over all rows and steps: 1.7e-2
    ckpt = _checkpoint_dir(tmp_path, cfg, sd)
    # five utterances of different lengths (one longer than 30 s), batch 2 -> three engine calls, the last one partial
    waves = [LM.synth_wave("chirp_short", seed=1), LM.synth_wave("noise_full", seed=2), LM.synth_wave("tones_long", seed=3),
             LM.synth_wave("chirp_short", seed=4)[:16000], LM.synth_wave("silence")]
    wav_dir = tmp_path / "wavs"
    wav_dir.mkdir()
    for i, w in enumerate(waves):
        audio.write_wav(str(wav_dir / f"utt-{i:04d}.wav"), w)
    paths, refs = audio.read_manifest(str(wav_dir))
    assert refs is None and len(paths) == 5
    pcm = [audio.load_audio(p) for p in paths]                        # what the pipeline reads (16-bit quantised)

    pipe = WhisperPipeline(ckpt, dtype="float32", max_batch=2, device="cuda:0", compact_every=8)
    ids = pipe.transcribe_files(paths)
    assert ids.shape == (5, 24) and ids.dtype == torch.int32 and not ids.is_cuda
    # oracle on the log-mel the GPU front-end produced: fp32 token ids are identical
    mel = pipe.frontend(pcm).cpu()
    ref = R.greedy(mel, sd, cfg)
    assert torch.equal(ids[:, :ref.shape[1]].long(), ref)
    assert bool((ids[:, ref.shape[1]:] == cfg["pad_token_id"]).all())
    # the front-end's log-mel is the oracle's (tests/test_logmel.py states and checks the tight bound; here: plumbing)
    want = np.stack([LM.log_mel(w) for w in pcm])
    assert np.abs(mel.numpy() - want).max() < 1e-2
    # same ids from waveforms and from features, whatever the batching
    assert torch.equal(pipe.transcribe_waveforms(pcm), ids)
    assert torch.equal(pipe.transcribe_features(mel), ids)
    assert pipe.transcribe_waveforms([]).shape == (0, 24)

    texts = pipe.decode(ids)
    assert len(texts) == 5 and all(isinstance(t, str) and t for t in texts)
    assert texts == pipe(pcm)
    first = ids[0, 2:].tolist()                                       # after <|startoftranscript|> and the forced token
    assert texts[0].startswith(pipe.detokenizer.decode(first[:3]))
    assert pipe.wer(texts, texts) == 0.0
    one_wrong = [texts[0] + " extra"] + texts[1:]
    n_words = sum(len(pipe.normalizer(t).split()) for t in texts)
    assert pipe.wer(one_wrong, texts) == pytest.approx(1.0 / n_words)
    pipe.close()


def _script(name):
    import importlib.util
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("examples_whisper_" + name, os.path.join(ROOT, "examples", "whisper", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_style_scripts_end_to_end(tmp_path, capsys):
    """build_encoder.py + build_decoder.py + run.py with the reference's flags (examples/whisper/): engine files and config.pkl in
    --engine_dir, runner classes + greedy_search over a directory of FLAC files; the printed transcriptions are those of the
    pipeline on the same files."""
    import flac_writer as FW
    from whisper_trtllm_b200.pipeline import WhisperPipeline
    cfg = synth.make_config("micro", max_length=16)
    sd = synth.make_weights(cfg, seed=4)
    ckpt = _checkpoint_dir(tmp_path, cfg, sd)
    engine_dir = str(tmp_path / "whisper_outputs")
    for name in ("build_encoder", "build_decoder"):
        _script(name).main(["--whisper", ckpt, "--engine_dir", engine_dir])
    data = tmp_path / "flac"
    data.mkdir()
    for i, kind in enumerate(("chirp_short", "silence", "chirp_short")):
        w = LM.synth_wave(kind, seed=10 + i)[:24000 + 8000 * i]
        pcm = np.clip(np.rint(w.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int64)
        (data / f"utt-{i}.flac").write_bytes(FW.encode(pcm, blocksize=4096, kind=("fixed", 2), partition_order=2))
    capsys.readouterr()
    _script("run").main(["--whisper", ckpt, "--engine_dir", engine_dir, "--dataset", str(data), "--batch", "2"])
    lines = capsys.readouterr().out.splitlines()
    assert lines[-1].startswith("B200 time:")
    printed = lines[-4:-1]
    pipe = WhisperPipeline(ckpt, dtype="float32", max_batch=2, device="cuda:0", compact_every=0)
    paths, _ = audio.read_manifest(str(data))
    want = pipe.decode(pipe.transcribe_files(paths))
    pipe.close()
    assert printed == want and all(printed)
