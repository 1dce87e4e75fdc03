"""Generate tests/golden/logmel.npz from the REAL reference feature extractor (build container only):

    python -m oracle.make_golden_logmel

Runs ``WhisperFeatureExtractor()(wave, sampling_rate=16000)`` of the vendored transformers tree on the seeded synthetic
waveforms of oracle/logmel_ref.synth_wave, records how far the restatement (oracle/logmel_ref.log_mel) is from it, and
stores sub-sampled reference features (every 13th frame) as the committed fixture."""
from __future__ import annotations

import json
import os

import numpy as np

from . import hf_reference as HF
from . import logmel_ref as LM

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
KINDS = ["noise_full", "chirp_short", "tones_long", "silence"]
FRAME_STRIDE = 13


def main():
    HF.import_reference()   # puts the vendored tree first on sys.path (and stubs its dependency pin check)
    from transformers.models.whisper.feature_extraction_whisper import WhisperFeatureExtractor
    fe = WhisperFeatureExtractor()
    arrays, meta = {}, {"frame_stride": FRAME_STRIDE, "cases": {}}
    assert np.abs(fe.mel_filters - LM.mel_filters()).max() < 1e-12
    for kind in KINDS:
        wave = LM.synth_wave(kind, seed=1)
        ref = fe(wave, sampling_rate=16000, return_tensors="np").input_features[0]
        mine = LM.log_mel(wave)
        assert ref.shape == (80, 3000) and mine.shape == (80, 3000)
        diff = float(np.abs(ref - mine).max())
        arrays[kind] = ref[:, ::FRAME_STRIDE].astype(np.float32)
        meta["cases"][kind] = {"samples": int(wave.size), "restatement_maxabs": diff, "ref_min": float(ref.min()), "ref_max": float(ref.max())}
        print(kind, "restatement max |diff| =", diff)
        assert diff < 1e-5
    np.savez_compressed(os.path.join(GOLDEN_DIR, "logmel.npz"), **arrays)
    with open(os.path.join(GOLDEN_DIR, "logmel.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
