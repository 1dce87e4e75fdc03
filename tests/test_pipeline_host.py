"""Host side of the scripts around the path (run.py:229-331, cal_wer.py:227-287 equivalents): audio files, manifests, tokenizer
files of a checkpoint directory, the --compare report.  No GPU: the engine itself is covered by the -m gpu tests."""
import json
import os
import wave

import numpy as np
import pytest

from whisper_trtllm_b200 import audio, pipeline
from whisper_trtllm_b200.text import bytes_to_unicode


def _tone(n=1600, f=440.0):
    return (0.5 * np.sin(2 * np.pi * f * np.arange(n) / audio.SAMPLING_RATE)).astype(np.float32)


def test_wav_roundtrip_is_within_half_an_lsb(tmp_path):
    x = _tone()
    p = str(tmp_path / "a.wav")
    audio.write_wav(p, x)
    y, rate = audio.read_wav(p)
    assert rate == 16000 and y.dtype == np.float32 and y.shape == x.shape
    assert np.abs(y - x).max() <= 0.5 / 32768 + 1e-7
    assert np.array_equal(audio.load_audio(p), y)
    # full-scale values are clipped, not wrapped
    audio.write_wav(p, np.array([1.0, -1.0, 2.0, -2.0], dtype=np.float32))
    assert audio.read_wav(p)[0].tolist() == [32767 / 32768, -1.0, 32767 / 32768, -1.0]


@pytest.mark.parametrize("width,dtype,scale", [(1, np.uint8, 128.0), (3, None, 8388608.0), (4, "<i4", 2147483648.0)])
def test_other_pcm_widths_and_stereo(tmp_path, width, dtype, scale):
    vals = np.array([0, 1, -1, 100, -100, int(scale) - 1, -int(scale)], dtype=np.int64)
    stereo = np.stack([vals, vals[::-1]], axis=1)                      # two channels, averaged on read
    if width == 1:
        raw = (stereo + 128).astype(np.uint8).tobytes()
    elif width == 3:
        raw = b"".join(int(v).to_bytes(3, "little", signed=True) for v in stereo.reshape(-1))
    else:
        raw = stereo.astype(dtype).tobytes()
    p = str(tmp_path / "w.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(width)
        w.setframerate(16000)
        w.writeframes(raw)
    y, _ = audio.read_wav(p)
    want = ((vals + vals[::-1]) / 2.0 / scale).astype(np.float32)
    assert np.allclose(y, want, rtol=0, atol=1e-7)


def test_wrong_rate_shape_and_format_are_errors(tmp_path):
    p = str(tmp_path / "r.wav")
    audio.write_wav(p, _tone(), rate=8000)
    with pytest.raises(ValueError, match="8000"):
        audio.load_audio(p)                                              # as the feature extractor: no silent resampling
    np.save(str(tmp_path / "m.npy"), np.zeros((2, 10), dtype=np.float32))
    with pytest.raises(ValueError, match="1-D"):
        audio.load_audio(str(tmp_path / "m.npy"))
    with pytest.raises(ValueError, match="unsupported"):
        audio.load_audio(str(tmp_path / "x.mp3"))
    np.save(str(tmp_path / "ok.npy"), _tone().astype(np.float64))
    assert audio.load_audio(str(tmp_path / "ok.npy")).dtype == np.float32


def test_manifests(tmp_path):
    d = tmp_path / "spk"
    d.mkdir()
    for uid in ("1-2-0001", "1-2-0000"):
        audio.write_wav(str(d / f"{uid}.wav"), _tone())
    np.save(str(d / "1-2-0002.npy"), _tone())
    # a bare directory: sorted audio files, no references
    paths, refs = audio.read_manifest(str(d))
    assert [os.path.basename(p) for p in paths] == ["1-2-0000.wav", "1-2-0001.wav", "1-2-0002.npy"] and refs is None
    # LibriSpeech transcript file (also found through its directory)
    (d / "1-2.trans.txt").write_text("1-2-0000 HELLO WORLD\n1-2-0002 TWENTY ONE\n")
    for target in (str(d / "1-2.trans.txt"), str(d)):
        paths, refs = audio.read_manifest(target)
        assert [os.path.basename(p) for p in paths] == ["1-2-0000.wav", "1-2-0002.npy"] and refs == ["HELLO WORLD", "TWENTY ONE"]
    (d / "bad.trans.txt").write_text("9-9-9999 NO AUDIO\n")
    with pytest.raises(FileNotFoundError):
        audio.read_manifest(str(d / "bad.trans.txt"))
    # TSV relative to the manifest, and jsonl
    (tmp_path / "m.tsv").write_text("spk/1-2-0000.wav\thello\nspk/1-2-0001.wav\tworld\n")
    paths, refs = audio.read_manifest(str(tmp_path / "m.tsv"))
    assert paths == [str(d / "1-2-0000.wav"), str(d / "1-2-0001.wav")] and refs == ["hello", "world"]
    (tmp_path / "m.jsonl").write_text(json.dumps({"audio": str(d / "1-2-0000.wav")}) + "\n")
    assert audio.read_manifest(str(tmp_path / "m.jsonl")) == ([str(d / "1-2-0000.wav")], None)
    (tmp_path / "mixed.tsv").write_text("a.wav\thello\nb.wav\n")
    with pytest.raises(ValueError, match="some"):
        audio.read_manifest(str(tmp_path / "mixed.tsv"))
    (tmp_path / "empty.tsv").write_text("\n")
    with pytest.raises(ValueError, match="empty"):
        audio.read_manifest(str(tmp_path / "empty.tsv"))
    e = tmp_path / "none"
    e.mkdir()
    with pytest.raises(FileNotFoundError):
        audio.read_manifest(str(e))


def test_librispeech_tree_and_mel_cache(tmp_path):
    import pickle

    import torch
    # LibriSpeech layout: <split>/<speaker>/<chapter>/<speaker>-<chapter>-<utt>.flac + <speaker>-<chapter>.trans.txt
    import flac_writer as FW
    want_paths, want_refs = [], []
    for spk, chap in (("1089", "134686"), ("1089", "134691"), ("121", "127105")):
        d = tmp_path / "test-clean" / spk / chap
        d.mkdir(parents=True)
        lines = []
        for u in range(2):
            uid = f"{spk}-{chap}-{u:04d}"
            pcm = np.rint(_tone(800 + 100 * u) * 32767).astype(np.int64)
            (d / f"{uid}.flac").write_bytes(FW.encode(pcm, blocksize=256))
            lines.append(f"{uid} TEXT OF {uid.replace('-', ' ')}")
        (d / f"{spk}-{chap}.trans.txt").write_text("\n".join(lines) + "\n")
    paths, refs = audio.read_manifest(str(tmp_path / "test-clean"))
    assert [os.path.basename(p) for p in paths] == ["1089-134686-0000.flac", "1089-134686-0001.flac", "1089-134691-0000.flac",
                                                    "1089-134691-0001.flac", "121-127105-0000.flac", "121-127105-0001.flac"]
    assert refs[0] == "TEXT OF 1089 134686 0000" and len(refs) == 6
    x = audio.load_audio(paths[1])
    assert x.shape == (900,) and np.abs(x - _tone(900)).max() < 1e-4
    # cal_wer.py's librispeech.cache: pickled (mel, text) pairs
    pairs = [(torch.full((80, 3000), float(i)), f"text {i}") for i in range(3)]
    with open(tmp_path / "librispeech.cache", "wb") as f:
        pickle.dump(pairs, f)
    mels, texts = audio.read_mel_cache(str(tmp_path / "librispeech.cache"))
    assert mels.shape == (3, 80, 3000) and mels.dtype == torch.float32 and float(mels[2, 5, 7]) == 2.0 and texts == ["text 0", "text 1", "text 2"]
    with open(tmp_path / "short.cache", "wb") as f:
        pickle.dump([(torch.zeros(80, 100), "x")], f)
    with pytest.raises(ValueError, match="3000"):
        audio.read_mel_cache(str(tmp_path / "short.cache"))


def test_batches():
    assert [list(b) for b in audio.batches(list(range(5)), 2)] == [[0, 1], [2, 3], [4]]
    assert list(audio.batches([], 4)) == []
    with pytest.raises(ValueError):
        list(audio.batches([1], 0))


def test_text_tools_from_a_checkpoint_directory(tmp_path):
    cfg = {"eos_token_id": 50256}
    # nothing there: ids only, normaliser with an empty spelling table
    detok, norm = pipeline.load_text_tools(str(tmp_path), cfg)
    assert detok is None and norm("The colour is twenty one") == "the colour is 21"
    b2u = bytes_to_unicode()
    vocab = {"".join(b2u[b] for b in " colour".encode()): 0, "".join(b2u[b] for b in " hi".encode()): 1,
             "<|endoftext|>": 50256, "<|startoftranscript|>": 50257}
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    (tmp_path / "normalizer.json").write_text(json.dumps({"colour": "color"}))
    detok, norm = pipeline.load_text_tools(str(tmp_path), cfg)
    assert detok.batch_decode([[50257, 1, 0, 50256, 50256]]) == [" hi colour"] and norm(" hi colour") == "hi color"
    assert pipeline.first_special_token_id(str(tmp_path), cfg) == 50256
    # `.en`: <|endoftext|> = 50256 sits in vocab.json, added_tokens.json starts at 50257 -> the block of specials starts at 50256
    (tmp_path / "added_tokens.json").write_text(json.dumps({"<|startoftranscript|>": 50257, "<|notimestamps|>": 50362}))
    assert pipeline.first_special_token_id(str(tmp_path), cfg) == 50256
    # multilingual: <|endoftext|> = eos = 50257 is itself the first added token
    (tmp_path / "added_tokens.json").write_text(json.dumps({"<|endoftext|>": 50257, "<|startoftranscript|>": 50258}))
    assert pipeline.first_special_token_id(str(tmp_path), {"eos_token_id": 50257}) == 50257


def test_compare_report():
    assert pipeline.compare_transcriptions(["a", "b", "c"], ["a", "x", "c"]) == [("b", "x")]
    with pytest.raises(ValueError):
        pipeline.compare_transcriptions(["a"], [])


def test_the_pipeline_has_no_cpu_path(tmp_path):
    import torch
    from oracle import synth
    from whisper_trtllm_b200 import WhisperB200Error, checkpoint
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = synth.make_config("micro")
    checkpoint.save_hf_checkpoint(str(tmp_path), cfg, synth.make_weights(cfg, seed=2))
    with pytest.raises(WhisperB200Error):
        pipeline.WhisperPipeline(str(tmp_path))


def test_cli_arguments():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "transcribe_cli", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "transcribe.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    a = cli.parse_arguments(["--whisper", "ckpt", "--audio", "dir", "--wer", "--batch", "8"])
    assert (a.whisper, a.audio, a.wer, a.compare, a.batch, a.dtype, a.compact_every) == ("ckpt", "dir", True, False, 8, "bfloat16", 32)


def test_sharded_transcription_prefetches_the_next_batch_and_keeps_order():
    """The scheduling around the engine, with the engine replaced by a stub: files of batch k + 1 are loaded while batch k is
    'on the GPU', results come back in input order, a short last batch is fine."""
    import threading

    import torch
    pipe = object.__new__(pipeline.WhisperPipeline)
    pipe.config = {"max_length": 6, "pad_token_id": 0}
    pipe.max_batch = 3
    events, lock = [], threading.Lock()

    def load(item):
        with lock:
            events.append(("load", item))
        return np.full(4, item, dtype=np.float32)

    def fake_transcribe(waves):
        with lock:
            events.append(("gpu", [int(w[0]) for w in waves]))
        return torch.tensor([[int(w[0])] * 6 for w in waves], dtype=torch.int32).reshape(len(waves), 6)

    pipe.transcribe_waveforms = fake_transcribe
    ids = pipe.transcribe_sharded(list(range(8)), load=load)
    assert ids[:, 0].tolist() == list(range(8)) and ids.shape == (8, 6)
    gpu = [e[1] for e in events if e[0] == "gpu"]
    assert gpu == [[0, 1, 2], [3, 4, 5], [6, 7]]
    # every file of batch 1 was handed to the pool before batch 0 went to the engine ... and nothing was loaded twice
    first_gpu = events.index(("gpu", [0, 1, 2]))
    assert sorted(e[1] for e in events if e[0] == "load") == list(range(8))
    assert {e[1] for e in events[:first_gpu] if e[0] == "load"} >= {0, 1, 2}
    # no loader: the items are waveforms already
    ids = pipe.transcribe_sharded([np.full(4, 9, dtype=np.float32)])
    assert ids.tolist() == [[9] * 6]
    assert pipe.transcribe_sharded([], load=load).shape[0] == 0


def _load_script(name):
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "whisper", name + ".py")
    spec = importlib.util.spec_from_file_location("examples_whisper_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_style_build_scripts(tmp_path, capsys):
    """examples/whisper/build_encoder.py / build_decoder.py: the reference's flags, engine files + config.pkl in --engine_dir
    (build_encoder.py:22-28, :42-45, :109; build_decoder.py:119), loadable by Session.from_serialized_engine."""
    import pickle

    import torch
    from oracle import synth
    from whisper_trtllm_b200 import checkpoint, run, runtime
    cfg = synth.make_config("micro")
    sd = synth.make_weights(cfg, seed=2)
    ckpt, out = str(tmp_path / "whisper-micro.en"), str(tmp_path / "whisper_outputs")
    checkpoint.save_hf_checkpoint(ckpt, cfg, sd)
    for name in ("build_encoder", "build_decoder"):
        script = _load_script(name)
        a = script.parse_arguments([])
        assert (a.whisper, a.engine_precision, a.log_level, a.engine_dir) == ("whisper-tiny.en", "float32", "error", "whisper_outputs")
        script.main(["--whisper", ckpt, "--engine_dir", out])
    assert sorted(os.listdir(out)) == ["WhisperDecoder.engine", "WhisperEncoder.engine", "config.pkl"]
    with open(os.path.join(out, "config.pkl"), "rb") as f:
        config = pickle.load(f)
    assert config["d_model"] == cfg["d_model"] and config["forced_decoder_ids"] == cfg["forced_decoder_ids"]
    with open(os.path.join(out, "WhisperEncoder.engine"), "rb") as f:
        kind, enc = runtime.deserialize_engine(f.read())
    assert kind == "WhisperEncoder" and torch.equal(enc.conv1.weight.data.cpu().reshape(-1), sd["model.encoder.conv1.weight"].reshape(-1))
    with open(os.path.join(out, "WhisperDecoder.engine"), "rb") as f:
        kind, dec = runtime.deserialize_engine(f.read())
    assert kind == "WhisperDecoder" and dec.dtype == torch.float32
    # bfloat16 engines
    _load_script("build_decoder").main(["--whisper", ckpt, "--engine_dir", out, "--engine_precision", "bfloat16"])
    with open(os.path.join(out, "WhisperDecoder.engine"), "rb") as f:
        assert runtime.deserialize_engine(f.read())[1].dtype == torch.bfloat16
    a = _load_script("run").parse_arguments(["--whisper", ckpt, "--compare"])
    assert (a.engine_dir, a.compare, a.dataset, a.batch) == ("whisper_outputs", True, "./librispeech_asr_dummy", 1)


def test_transcribe_cli_flow_with_a_stub_engine(tmp_path, monkeypatch, capsys):
    """examples/transcribe.py end to end on the host: manifest / HF-style inputs -> (stub) pipeline -> printed transcripts, --out
    JSON lines, --wer.  The stub stands where the GPU pipeline is; everything around it is the real code."""
    import importlib.util
    import pickle

    import torch
    spec = importlib.util.spec_from_file_location(
        "transcribe_cli2", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "transcribe.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)

    class FakePipeline:
        made = []

        def __init__(self, checkpoint_dir, dtype, max_batch, compact_every):
            self.args = (checkpoint_dir, dtype, max_batch, compact_every)
            FakePipeline.made.append(self)
            from whisper_trtllm_b200.text import EnglishTextNormalizer
            self.normalizer = EnglishTextNormalizer()

        def transcribe_files(self, paths):
            return torch.tensor([[len(audio.load_audio(p))] for p in paths], dtype=torch.int32)

        def transcribe_sharded(self, items, load=None, features=False):
            return torch.tensor([[int(m[0, 0])] for m in items], dtype=torch.int32) if features else \
                torch.tensor([[len(w)] for w in items], dtype=torch.int32)

        def decode(self, ids):
            return [f"utterance of {int(r[0])} samples" for r in ids]

        def wer(self, hyp, ref):
            from whisper_trtllm_b200.text import wer
            return wer([self.normalizer(t) for t in ref], [self.normalizer(t) for t in hyp])

        def close(self):
            self.closed = True

    from whisper_trtllm_b200 import pipeline as pl
    monkeypatch.setattr(pl, "WhisperPipeline", FakePipeline)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *_: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *_: None)
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)

    d = tmp_path / "wavs"
    d.mkdir()
    audio.write_wav(str(d / "a.wav"), _tone(800))
    audio.write_wav(str(d / "b.wav"), _tone(1200))
    (tmp_path / "m.tsv").write_text("wavs/a.wav\tutterance of 800 samples\nwavs/b.wav\tutterance of twelve hundred samples\n")
    out = tmp_path / "out.jsonl"
    cli.main(["--whisper", "ckpt", "--audio", str(tmp_path / "m.tsv"), "--wer", "--out", str(out), "--batch", "4", "--dtype", "float32"])
    printed = capsys.readouterr().out.splitlines()
    assert printed[:2] == ["a.wav\tutterance of 800 samples", "b.wav\tutterance of 1200 samples"]
    assert printed[2] == "WER: 0.00 %"                       # "twelve hundred" and "1200" normalise to the same words
    rows = [json.loads(ln) for ln in out.read_text().splitlines()]
    assert rows[1] == {"audio": str(d / "b.wav"), "text": "utterance of 1200 samples", "reference": "utterance of twelve hundred samples"}
    assert FakePipeline.made[-1].args == ("ckpt", "float32", 4, 32) and FakePipeline.made[-1].closed
    # a directory without references: --wer is refused before any work is done
    with pytest.raises(SystemExit):
        cli.main(["--whisper", "ckpt", "--audio", str(d), "--wer"])
    # cal_wer.py's librispeech.cache goes through the feature path
    with open(tmp_path / "librispeech.cache", "wb") as f:
        pickle.dump([(torch.full((80, 3000), 5.0), "utterance of five samples"), (torch.full((80, 3000), 7.0), "nothing alike")], f)
    cli.main(["--whisper", "ckpt", "--audio", str(tmp_path / "librispeech.cache"), "--wer"])
    printed = capsys.readouterr().out.splitlines()
    assert printed[0] == "librispeech.cache[0]\tutterance of 5 samples" and printed[-1] == "WER: 66.67 %"
