"""GPU log-mel front-end: the step BEFORE the hot path.  The reference computes
``hf_processor(sample["array"], sampling_rate=...).input_features`` with numpy on one CPU core per utterance
(run.py:267; feature_extraction_whisper.py:96-109 over audio_utils.spectrogram); at > 1000x real time that would
dominate the wall clock.  Here the framing, DFT (as an fp32 GEMM), mel projection and log/clamp run in
libwhisper_b200 (csrc/frontend.cu); this module only builds the constant tables and pads / trims the waveforms.
"""
from __future__ import annotations

from typing import List, Sequence, Union

import numpy as np
import torch

from . import _abi
from ._abi import byref, c_size_t, ptr, stream_handle

SAMPLING_RATE, N_FFT, HOP, N_MELS = 16000, 400, 160, 80
N_SAMPLES, N_FRAMES, N_BINS = 480000, 3000, 201
SPEC_LD, POW_LD = 408, 208


def _mel_filters() -> np.ndarray:
    """[201, 80] slaney-scale, slaney-normalised triangular filters, 0-8000 Hz (audio_utils.py:115-190)."""
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)
    fft_freqs = np.linspace(0, SAMPLING_RATE // 2, N_BINS)
    f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(8000.0), N_MELS + 2))
    diff = np.diff(f)
    slopes = f[None, :] - fft_freqs[:, None]
    fb = np.maximum(0.0, np.minimum(-slopes[:, :-2] / diff[:-1], slopes[:, 2:] / diff[1:]))
    return fb * (2.0 / (f[2:N_MELS + 2] - f[:N_MELS]))[None, :]


class LogMelFrontend:
    """``frontend(waves) -> input_features fp32 [B, 80, 3000]`` on the device, the tensor ``WhisperEncoder`` consumes."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise _abi.WhisperB200Error(-101, "a CUDA device is required (no CPU fallback)")
        _abi.load()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        k = np.arange(N_BINS, dtype=np.float64)[:, None]
        j = np.arange(N_FFT, dtype=np.float64)[None, :]
        ang = 2.0 * np.pi * k * j / N_FFT
        basis = np.zeros((SPEC_LD, N_FFT), dtype=np.float64)
        basis[:N_BINS] = np.cos(ang)
        basis[N_BINS:2 * N_BINS] = -np.sin(ang)
        filt = np.zeros((N_MELS, POW_LD), dtype=np.float64)
        filt[:, :N_BINS] = _mel_filters().T
        window = np.hanning(N_FFT + 1)[:-1]                       # periodic Hann (audio_utils.py:237-247)
        to = lambda a: torch.from_numpy(a.astype(np.float32)).to(self.device).contiguous()
        self.window, self.basis, self.filters = to(window), to(basis), to(filt)
        self._ws = None

    @staticmethod
    def pad_or_trim(waves: Union[np.ndarray, torch.Tensor, Sequence]) -> torch.Tensor:
        """Zero-pad / truncate every waveform to 30 s (padding='max_length', truncation=True; FE:213-235) -> fp32 [B, 480000]."""
        if isinstance(waves, torch.Tensor) and waves.dim() == 2 and waves.shape[1] == N_SAMPLES:
            return waves.to(torch.float32)
        if isinstance(waves, (np.ndarray, torch.Tensor)) and getattr(waves, "ndim", 1) == 1:
            waves = [waves]
        out = torch.zeros(len(waves), N_SAMPLES, dtype=torch.float32)
        for i, w in enumerate(waves):
            w = torch.as_tensor(np.asarray(w, dtype=np.float32) if not isinstance(w, torch.Tensor) else w).reshape(-1).to(torch.float32).cpu()
            n = min(w.numel(), N_SAMPLES)
            out[i, :n] = w[:n]
        return out

    @torch.no_grad()
    def __call__(self, waves, stream=None) -> torch.Tensor:
        pcm = self.pad_or_trim(waves)
        if not pcm.is_cuda:
            pcm = pcm.pin_memory().to(self.device, non_blocking=True)
        pcm = pcm.contiguous()
        B = pcm.shape[0]
        nbytes = c_size_t()
        _abi.call("wb_log_mel_workspace_bytes", B, byref(nbytes))
        if self._ws is None or self._ws.numel() < nbytes.value:
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        out = torch.empty(B, N_MELS, N_FRAMES, dtype=torch.float32, device=self.device)
        _abi.call("wb_log_mel", ptr(pcm), B, ptr(self.window), ptr(self.basis), ptr(self.filters), ptr(self._ws),
                  c_size_t(self._ws.numel()), ptr(out), stream_handle(stream))
        return out
