"""`WhisperProcessor` stand-in for the two calls the reference's scripts make on it (run.py:239, :267, :287; cal_wer.py:237, :275):

    hf_processor = WhisperProcessor.from_pretrained(args.whisper)
    input_features = hf_processor(sample["array"], sampling_rate=sample["sampling_rate"], return_tensors="pt").input_features
    transcription = hf_processor.batch_decode(predicted_ids, skip_special_tokens=True)

so that the reference-side edit is the import line.  The features are computed by the GPU front-end (`LogMelFrontend`, csrc/frontend.cu)
and come back as a CUDA tensor (the reference moves them there on the next line anyway, run.py:268); ids -> text is
`text.WhisperDetokenizer`.  Argument checking follows `WhisperFeatureExtractor.__call__` (feature_extraction_whisper.py:136-260):
a sampling rate other than 16 kHz is a ValueError, more than one channel is a ValueError; options of the reference that would
change the result and are not implemented here (`do_normalize`, other padding / truncation modes, attention masks, non-"pt"
tensors) raise instead of being ignored.  Host plumbing only; there is no CPU feature path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .audio import SAMPLING_RATE

N_SAMPLES = 30 * SAMPLING_RATE


class BatchFeature(dict):
    """Dictionary with attribute access, like transformers' BatchFeature (`.input_features`)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)


def as_batch(raw_speech) -> List[np.ndarray]:
    """The reference's input conventions (feature_extraction_whisper.py:208-226): one waveform (1-D array / list of floats) or a
    batch (2-D array, or a list / tuple of arrays / lists) -> list of 1-D float32 arrays.  More than two dimensions = more than
    one channel = ValueError."""
    try:
        import torch
        if isinstance(raw_speech, torch.Tensor):
            raw_speech = raw_speech.detach().cpu().numpy()
    except ImportError:                                    # pragma: no cover
        pass
    if isinstance(raw_speech, np.ndarray):
        if raw_speech.ndim > 2:
            raise ValueError("Only mono-channel audio is supported for input to WhisperFeatureExtractor")
        rows = list(raw_speech) if raw_speech.ndim == 2 else [raw_speech]
    elif isinstance(raw_speech, (list, tuple)) and len(raw_speech) and isinstance(raw_speech[0], (np.ndarray, list, tuple)):
        rows = list(raw_speech)
    else:
        rows = [raw_speech]
    out = []
    for r in rows:
        a = np.asarray(r, dtype=np.float32)
        if a.ndim != 1:
            raise ValueError("Only mono-channel audio is supported for input to WhisperFeatureExtractor")
        out.append(a)
    return out


class WhisperFeatureExtractor:
    sampling_rate = SAMPLING_RATE
    n_samples = N_SAMPLES
    feature_size = 80
    hop_length = 160
    n_fft = 400

    def __init__(self, frontend=None, device=None):
        self._frontend = frontend
        self._device = device

    @property
    def frontend(self):
        if self._frontend is None:                         # built on first use: needs the GPU and libwhisper_b200.so
            from .frontend import LogMelFrontend
            self._frontend = LogMelFrontend(self._device)
        return self._frontend

    def __call__(self, raw_speech, truncation: bool = True, pad_to_multiple_of: Optional[int] = None, return_tensors: Optional[str] = "pt",
                 return_attention_mask: Optional[bool] = None, padding: Optional[str] = "max_length", max_length: Optional[int] = None,
                 sampling_rate: Optional[int] = None, do_normalize: Optional[bool] = None, **kwargs) -> BatchFeature:
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(
                f"The model corresponding to this feature extractor: {self.__class__.__name__} was trained using a sampling rate of "
                f"{self.sampling_rate}. Please make sure that the provided `raw_speech` input was sampled with {self.sampling_rate} "
                f"and not {sampling_rate}.")
        unsupported = {"truncation": truncation is not True, "pad_to_multiple_of": pad_to_multiple_of is not None,
                       "return_attention_mask": bool(return_attention_mask), "padding": padding != "max_length",
                       "max_length": max_length not in (None, self.n_samples), "do_normalize": bool(do_normalize),
                       "return_tensors": return_tensors != "pt", **{k: True for k in kwargs}}
        bad = sorted(k for k, v in unsupported.items() if v)
        if bad:
            raise NotImplementedError(f"WhisperFeatureExtractor option(s) {bad} are not implemented by the GPU front-end: 30 s windows "
                                      "(pad / cut), no normalisation, torch tensors on the device")
        return BatchFeature({"input_features": self.frontend(as_batch(raw_speech))})


class WhisperProcessor:
    def __init__(self, feature_extractor: WhisperFeatureExtractor, tokenizer=None):
        self.feature_extractor = feature_extractor
        self.tokenizer = tokenizer                         # text.WhisperDetokenizer (ids -> text only)

    @classmethod
    def from_pretrained(cls, checkpoint_dir: str, device=None) -> "WhisperProcessor":
        """Tokenizer files (`vocab.json`, `added_tokens.json`, `tokenizer_config.json`) from the checkpoint directory; the
        feature extractor has no files of its own (its constants are Whisper's: 16 kHz, n_fft 400, hop 160, 80 mels, 30 s)."""
        from .checkpoint import load_config
        from .pipeline import load_text_tools
        tokenizer, _ = load_text_tools(checkpoint_dir, load_config(checkpoint_dir))
        return cls(WhisperFeatureExtractor(device=device), tokenizer)

    def __call__(self, audio=None, **kwargs) -> BatchFeature:
        if audio is None:
            raise ValueError("You need to specify an `audio` input to process.")
        return self.feature_extractor(audio, **kwargs)

    def _need_tokenizer(self):
        if self.tokenizer is None:
            raise FileNotFoundError("the checkpoint directory has no vocab.json: token ids cannot be turned into text")
        return self.tokenizer

    def batch_decode(self, sequences, skip_special_tokens: bool = False, **kwargs) -> List[str]:
        return self._need_tokenizer().batch_decode(sequences, skip_special_tokens, kwargs.get("clean_up_tokenization_spaces"))

    def decode(self, token_ids: Sequence[int], skip_special_tokens: bool = False, **kwargs) -> str:
        ids = token_ids.tolist() if hasattr(token_ids, "tolist") else token_ids
        return self._need_tokenizer().decode(ids, skip_special_tokens, kwargs.get("clean_up_tokenization_spaces"))
