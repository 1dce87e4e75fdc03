"""CPU restatement of the reference's log-mel front-end (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows ``WhisperFeatureExtractor`` of the vendored transformers tree (``FE`` =
/root/reference/transformers/src/transformers/models/whisper/feature_extraction_whisper.py, ``AU`` =
/root/reference/transformers/src/transformers/audio_utils.py):

  pad / truncate to 30 s = 480 000 samples with zeros                       FE:213-235 (padding="max_length", truncation)
  reflect-pad 200 samples each side, frames of 400 with hop 160              AU:379-386
  periodic Hann window (np.hanning(401)[:-1])                                AU:237-247, FE:101
  rfft -> complex64 -> |.|^2 in float64                                      AU:389-409
  mel filter bank: 80 slaney-scale, slaney-normalised triangles, 0-8000 Hz   AU:115-190, FE:86-94
  max(1e-10, filters^T @ power) -> log10 -> float32                          AU:413-433
  drop the last frame, clamp to (max - 8), (x + 4) / 4                       FE:106-109

Pinned against the real reference by oracle/make_golden_logmel.py (tests/golden/logmel.npz).
"""
from __future__ import annotations

import numpy as np

SAMPLING_RATE, N_FFT, HOP, N_MELS, CHUNK_SECONDS = 16000, 400, 160, 80, 30
N_SAMPLES = SAMPLING_RATE * CHUNK_SECONDS          # 480000
N_FRAMES = N_SAMPLES // HOP                        # 3000
N_BINS = N_FFT // 2 + 1                            # 201


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    log_region = f >= 1000.0
    with np.errstate(divide="ignore"):
        mels = np.where(log_region, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * (27.0 / np.log(6.4)), mels)
    return mels


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), f)


def mel_filters() -> np.ndarray:
    """[201, 80] float64 (AU:115-190 with norm='slaney', mel_scale='slaney', 0..8000 Hz, 16 kHz)."""
    fft_freqs = np.linspace(0, SAMPLING_RATE // 2, N_BINS)
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(8000.0), N_MELS + 2)
    f = _mel_to_hz_slaney(mel_pts)
    diff = np.diff(f)
    slopes = f[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (f[2:N_MELS + 2] - f[:N_MELS]))[None, :]
    return fb


def hann_window() -> np.ndarray:
    return np.hanning(N_FFT + 1)[:-1]


def pad_or_trim(wave: np.ndarray) -> np.ndarray:
    wave = np.asarray(wave, dtype=np.float32).reshape(-1)
    if wave.size >= N_SAMPLES:
        return wave[:N_SAMPLES]
    return np.concatenate([wave, np.zeros(N_SAMPLES - wave.size, dtype=np.float32)])


def log_mel(wave: np.ndarray) -> np.ndarray:
    """One utterance: raw 16 kHz mono PCM (any length) -> input_features float32 [80, 3000]."""
    x = pad_or_trim(wave)
    x = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect").astype(np.float64)
    n_frames = 1 + (x.size - N_FFT) // HOP       # 3001
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    frames = x[idx] * hann_window()[None, :]
    spec = np.fft.rfft(frames, axis=1).astype(np.complex64)           # the reference stores complex64 (AU:389)
    power = np.abs(spec, dtype=np.float64) ** 2.0
    mel = np.maximum(1e-10, mel_filters().T @ power.T)                # [80, 3001]
    log_spec = np.log10(mel).astype(np.float32)
    log_spec = log_spec[:, :-1]
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return ((log_spec + 4.0) / 4.0).astype(np.float32)


def synth_wave(kind: str, seed: int = 0) -> np.ndarray:
    """Seeded synthetic 16 kHz waveforms (bit-reproducible: integer generator + exact fp32 steps + numpy sin in float64)."""
    rng = np.random.RandomState(seed)
    if kind == "noise_full":      # 30 s of noise with a slow amplitude envelope
        n = N_SAMPLES
        w = (rng.randint(0, 1 << 16, n).astype(np.float32) / 32768.0 - 1.0) * 0.3
        return (w * (0.2 + 0.8 * np.abs(np.sin(np.arange(n, dtype=np.float64) * 1e-4))).astype(np.float32)).astype(np.float32)
    if kind == "chirp_short":     # 7.3 s chirp (needs zero padding)
        n = int(7.3 * SAMPLING_RATE)
        t = np.arange(n, dtype=np.float64) / SAMPLING_RATE
        return (0.5 * np.sin(2 * np.pi * (200.0 * t + 400.0 * t * t))).astype(np.float32)
    if kind == "tones_long":      # 33 s (needs truncation), three tones + tiny noise
        n = 33 * SAMPLING_RATE
        t = np.arange(n, dtype=np.float64) / SAMPLING_RATE
        w = 0.3 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 1234.5 * t) + 0.1 * np.sin(2 * np.pi * 6000 * t)
        return (w + (rng.randint(0, 1 << 16, n) / 32768.0 - 1.0) * 1e-3).astype(np.float32)
    if kind == "silence":
        return np.zeros(N_SAMPLES // 2, dtype=np.float32)
    raise ValueError(kind)
