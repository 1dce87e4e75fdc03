"""Audio files and evaluation manifests for the scripts around the path (host-side I/O only, no arithmetic on the path).

The reference's scripts take their audio from an HF `datasets` table (``ds[i]["audio"]["array"]`` at 16 kHz, run.py:266-267) or
from a pickled list of (mel, text) pairs (cal_wer.py:248-249); neither the table nor its decoders (soundfile / torchcodec)
are in this image, so the harness reads what the standard library can: PCM ``.wav`` files and ``.npy`` arrays, listed either
by a directory, a LibriSpeech-style ``*.trans.txt`` (``<utterance-id> <TEXT>`` per line, audio next to it), a TSV
(``<path>\\t<text>``) or a JSON-lines file (``{"audio": path, "text": ...}``).
"""
from __future__ import annotations

import json
import os
import wave
from typing import List, Optional, Sequence, Tuple

import numpy as np

SAMPLING_RATE = 16000          # feature_extraction_whisper.py:60-66 (sampling_rate=16000); other rates are rejected (:195-201)
AUDIO_EXTENSIONS = (".wav", ".npy")


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
    if ext == ".wav":
        x, rate = read_wav(path)
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

    * a directory: every .wav / .npy in it (sorted); references from the ``*.trans.txt`` files in it when there are any;
    * ``*.trans.txt``: LibriSpeech transcript file, ``<id> <TEXT>``, audio = ``<dir>/<id>.wav|.npy``;
    * ``*.jsonl``: one object per line with ``audio`` (path, relative to the file) and optionally ``text``;
    * anything else: TSV ``<path>[\\t<text>]``; a line without a tab is an audio path without a reference.
    """
    if os.path.isdir(path):
        trans = sorted(f for f in os.listdir(path) if f.endswith(".trans.txt"))
        if trans:
            audio: List[str] = []
            texts: List[str] = []
            for t in trans:
                a, r = read_manifest(os.path.join(path, t))
                audio += a
                texts += r or []
            return audio, texts
        files = sorted(f for f in os.listdir(path) if f.lower().endswith(AUDIO_EXTENSIONS))
        if not files:
            raise FileNotFoundError(f"no {' / '.join(AUDIO_EXTENSIONS)} files under {path}")
        return [os.path.join(path, f) for f in files], None
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


def batches(items: Sequence, size: int):
    """Consecutive slices of at most `size` items (the last one may be shorter)."""
    if size <= 0:
        raise ValueError("batch size must be positive")
    for i in range(0, len(items), size):
        yield items[i:i + size]
