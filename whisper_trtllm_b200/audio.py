"""Audio files and evaluation manifests for the scripts around the path (host-side I/O only, no arithmetic on the path).

The reference's scripts take their audio from an HF `datasets` table (``ds[i]["audio"]["array"]`` at 16 kHz, run.py:266-267) or
from a pickled list of (mel, text) pairs (cal_wer.py:248-249); neither the table nor its decoders (soundfile / torchcodec)
are in this image, so the harness brings its own readers: PCM ``.wav`` (standard library), ``.npy`` arrays and ``.flac`` — the
format of LibriSpeech and of the reference's bundled ``librispeech_asr_dummy`` table — through the repo's own FLAC decoder
(``libwb_audio.so``, csrc/audio/flac_decode.c, include/wb_audio.h; frame CRCs and the stream's MD5 signature are verified).
Utterances are listed by a directory, a LibriSpeech-style ``*.trans.txt`` (``<utterance-id> <TEXT>`` per line, audio next to
it), a TSV (``<path>\\t<text>``), a JSON-lines file (``{"audio": path, "text": ...}``) or an HF ``datasets`` directory saved with
``save_to_disk`` (``read_hf_dataset``: the audio column is read undecoded and decoded here).
"""
from __future__ import annotations

import ctypes
import json
import os
import wave
from typing import List, Optional, Sequence, Tuple

import numpy as np

SAMPLING_RATE = 16000          # feature_extraction_whisper.py:60-66 (sampling_rate=16000); other rates are rejected (:195-201)
AUDIO_EXTENSIONS = (".wav", ".flac", ".npy")
AUDIO_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libwb_audio.so")


class AudioDecodeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libwb_audio error {code}: {msg}")
        self.code = code


class wb_flac_info(ctypes.Structure):
    _fields_ = [("sample_rate", ctypes.c_uint32), ("channels", ctypes.c_uint32), ("bits_per_sample", ctypes.c_uint32),
                ("min_blocksize", ctypes.c_uint32), ("max_blocksize", ctypes.c_uint32), ("total_samples", ctypes.c_uint64),
                ("audio_offset", ctypes.c_uint64), ("md5", ctypes.c_uint8 * 16)]


# name -> (restype, argtypes): every symbol include/wb_audio.h declares
AUDIO_SIGNATURES = {
    "wb_audio_last_error": (ctypes.c_char_p, []),
    "wb_audio_version": (ctypes.c_int, []),
    "wb_flac_read_info": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(wb_flac_info)]),
    "wb_flac_decode_i32": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_uint64,
                                          ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]),
    "wb_md5": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]),
}
_audio_lib = None


def load_audio_lib():
    """libwb_audio.so (host-only, built by csrc/Makefile).  No fallback decoder: a missing library is an error."""
    global _audio_lib
    if _audio_lib is None:
        if not os.path.exists(AUDIO_LIB_PATH):
            raise AudioDecodeError(-100, f"{AUDIO_LIB_PATH} is missing: build it with `make -C whisper_trtllm_b200/csrc`")
        lib = ctypes.CDLL(AUDIO_LIB_PATH)
        for name, (res, args) in AUDIO_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _audio_lib = lib
    return _audio_lib


def _audio_call(name: str, *args):
    lib = load_audio_lib()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise AudioDecodeError(rc, lib.wb_audio_last_error().decode("utf-8", "replace"))


def flac_info(data: bytes) -> dict:
    info = wb_flac_info()
    _audio_call("wb_flac_read_info", data, len(data), ctypes.byref(info))
    return {"sample_rate": info.sample_rate, "channels": info.channels, "bits_per_sample": info.bits_per_sample,
            "total_samples": info.total_samples, "min_blocksize": info.min_blocksize, "max_blocksize": info.max_blocksize,
            "md5": bytes(info.md5).hex()}


def decode_flac_pcm(data: bytes, verify_md5: bool = True) -> Tuple[np.ndarray, dict]:
    """FLAC bytes -> (int32 PCM [samples, channels] with the stream's native sample values, stream info).  Every frame's CRC is
    checked, and (``verify_md5``) the MD5 of the decoded PCM against the signature in STREAMINFO when the stream carries one."""
    info = flac_info(data)
    capacity = info["total_samples"]
    if capacity == 0:            # length not recorded in STREAMINFO: count first (out = NULL)
        counted = ctypes.c_uint64()
        _audio_call("wb_flac_decode_i32", data, len(data), None, ctypes.c_uint64(0), ctypes.byref(counted), 0)
        capacity = counted.value
    out = np.empty((max(capacity, 1), info["channels"]), dtype=np.int32)
    n = ctypes.c_uint64()
    _audio_call("wb_flac_decode_i32", data, len(data), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(capacity), ctypes.byref(n),
                1 if verify_md5 else 0)
    return out[:n.value], info


def decode_flac(data: bytes, verify_md5: bool = True) -> Tuple[np.ndarray, int]:
    """FLAC bytes -> (float32 mono waveform in [-1, 1), sampling rate); channels are averaged, as `datasets` does for mono."""
    pcm, info = decode_flac_pcm(data, verify_md5)
    x = pcm.astype(np.float64).mean(axis=1) / float(1 << (info["bits_per_sample"] - 1))
    return x.astype(np.float32), info["sample_rate"]


def read_wav(path: str) -> Tuple[np.ndarray, int]:
    """PCM .wav (8 / 16 / 24 / 32-bit integer) -> (float32 mono waveform in [-1, 1), sampling rate).  Channels are averaged."""
    with wave.open(path, "rb") as w:
        n_ch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width == 1:                                   # unsigned 8-bit
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 3:                                 # 24-bit little endian -> sign-extended int32
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    else:
        raise ValueError(f"{path}: unsupported sample width {width}")
    if n_ch > 1:
        x = x.reshape(-1, n_ch).mean(axis=1)
    return np.ascontiguousarray(x, dtype=np.float32), rate


def write_wav(path: str, wave_f32: np.ndarray, rate: int = SAMPLING_RATE) -> None:
    """float waveform in [-1, 1] -> 16-bit PCM mono .wav (round to nearest, clipped)."""
    pcm = np.clip(np.rint(np.asarray(wave_f32, dtype=np.float64) * 32768.0), -32768, 32767).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(pcm.tobytes())


def load_audio(path: str, sampling_rate: int = SAMPLING_RATE) -> np.ndarray:
    """One utterance as a float32 waveform at 16 kHz.  A file at another rate is an error, as it is in the reference's
    feature extractor (no silent resampling)."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        x = np.load(path)
        if x.ndim != 1:
            raise ValueError(f"{path}: expected a 1-D waveform, got shape {x.shape}")
        return np.ascontiguousarray(x, dtype=np.float32)
    if ext in (".wav", ".flac"):
        if ext == ".wav":
            x, rate = read_wav(path)
        else:
            with open(path, "rb") as f:
                x, rate = decode_flac(f.read())
        if rate != sampling_rate:
            raise ValueError(f"{path}: sampling rate {rate} Hz, the model was trained on {sampling_rate} Hz audio — resample first")
        return x
    raise ValueError(f"{path}: unsupported audio format {ext!r} (supported: {', '.join(AUDIO_EXTENSIONS)})")


def _find_audio(stem: str) -> Optional[str]:
    for ext in AUDIO_EXTENSIONS:
        if os.path.exists(stem + ext):
            return stem + ext
    return None


def read_manifest(path: str) -> Tuple[List[str], Optional[List[str]]]:
    """-> (audio paths, reference texts or None).

    * a directory: every .wav / .flac / .npy in it and below it (sorted by path); when there are ``*.trans.txt`` files anywhere
      below it (a LibriSpeech split: ``test-clean/<speaker>/<chapter>/``), the utterances and references they list;
    * ``*.trans.txt``: LibriSpeech transcript file, ``<id> <TEXT>``, audio = ``<dir>/<id>.wav|.npy``;
    * ``*.jsonl``: one object per line with ``audio`` (path, relative to the file) and optionally ``text``;
    * anything else: TSV ``<path>[\\t<text>]``; a line without a tab is an audio path without a reference.
    """
    if os.path.isdir(path):            # the directory and everything below it (LibriSpeech: <split>/<speaker>/<chapter>/...)
        trans, files = [], []
        for root, dirs, names in os.walk(path):
            dirs.sort()
            trans += [os.path.join(root, f) for f in sorted(names) if f.endswith(".trans.txt")]
            files += [os.path.join(root, f) for f in sorted(names) if f.lower().endswith(AUDIO_EXTENSIONS)]
        if trans:
            audio: List[str] = []
            texts: List[str] = []
            for t in trans:
                a, r = read_manifest(t)
                audio += a
                texts += r or []
            return audio, texts
        if not files:
            raise FileNotFoundError(f"no {' / '.join(AUDIO_EXTENSIONS)} files under {path}")
        return files, None
    base = os.path.dirname(os.path.abspath(path))
    audio, texts, have_text = [], [], []
    with open(path, encoding="utf-8") as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    for ln in lines:
        if path.endswith(".trans.txt"):
            uid, _, text = ln.partition(" ")
            p = _find_audio(os.path.join(base, uid))
            if p is None:
                raise FileNotFoundError(f"{path}: no audio file for utterance {uid!r}")
            text_present = True
        elif path.endswith(".jsonl"):
            o = json.loads(ln)
            p, text, text_present = o["audio"], o.get("text", ""), "text" in o
        else:
            p, tab, text = ln.partition("\t")
            text_present = bool(tab)
        audio.append(p if os.path.isabs(p) else os.path.join(base, p))
        texts.append(text)
        have_text.append(text_present)
    if not audio:
        raise ValueError(f"{path}: empty manifest")
    if any(have_text) and not all(have_text):
        raise ValueError(f"{path}: some utterances have a reference text and some do not")
    return audio, (texts if all(have_text) else None)


def is_hf_dataset(path: str) -> bool:
    """A directory written by `datasets`' ``save_to_disk`` (it holds state.json / dataset_info.json, or dataset_dict.json)."""
    return os.path.isdir(path) and any(os.path.exists(os.path.join(path, f)) for f in ("state.json", "dataset_info.json", "dataset_dict.json"))


def read_hf_dataset(path: str, audio_column: str = "audio", text_column: str = "text", sampling_rate: int = SAMPLING_RATE):
    """An HF `datasets` directory written by ``save_to_disk`` (what run.py:241-247 loads as ./librispeech_asr_dummy) ->
    (list of float32 waveforms, reference texts or None).  The audio column is read UNDECODED (``Audio(decode=False)``: the
    table stores the file bytes) and decoded here, so neither soundfile nor torchcodec is needed."""
    import datasets
    ds = datasets.load_from_disk(path)
    if isinstance(ds, datasets.DatasetDict):
        ds = ds[next(iter(ds))]
    if audio_column not in ds.column_names:
        raise KeyError(f"{path}: no column {audio_column!r} (columns: {ds.column_names})")
    if isinstance(ds.features[audio_column], datasets.Audio):
        ds = ds.cast_column(audio_column, datasets.Audio(decode=False))
    waves = []
    for i, item in enumerate(ds[audio_column]):
        data, name = item.get("bytes"), item.get("path") or ""
        if data is None:
            waves.append(load_audio(name, sampling_rate))
            continue
        if data[:4] == b"fLaC" or data[:3] == b"ID3":
            x, rate = decode_flac(data)
        elif data[:4] == b"RIFF":
            import io
            x, rate = read_wav(io.BytesIO(data))
        else:
            raise ValueError(f"{path}: row {i} ({name}) is neither FLAC nor WAV")
        if rate != sampling_rate:
            raise ValueError(f"{path}: row {i} ({name}) has sampling rate {rate} Hz, expected {sampling_rate} Hz")
        waves.append(x)
    texts = list(ds[text_column]) if text_column in ds.column_names else None
    return waves, texts


def read_mel_cache(path: str):
    """The reference's ``librispeech.cache`` (get_LibriSpeech.py:32-41, read by cal_wer.py:248-249): a pickled list of
    (log-mel [80, 3000], text) pairs -> (fp32 tensor [n, 80, 3000], texts).  A pickle runs code when loaded: only open files you
    made yourself, exactly as with the reference's script."""
    import pickle

    import torch
    with open(path, "rb") as f:
        pairs = pickle.load(f)
    if not pairs:
        raise ValueError(f"{path}: empty cache")
    mels = torch.stack([torch.as_tensor(m, dtype=torch.float32).reshape(80, -1) for m, _ in pairs])
    if mels.shape[-1] != 3000:
        raise ValueError(f"{path}: log-mel windows of {mels.shape[-1]} frames, expected 3000 (30 s)")
    return mels.contiguous(), [t for _, t in pairs]


def batches(items: Sequence, size: int):
    """Consecutive slices of at most `size` items (the last one may be shorter)."""
    if size <= 0:
        raise ValueError("batch size must be positive")
    for i in range(0, len(items), size):
        yield items[i:i + size]
