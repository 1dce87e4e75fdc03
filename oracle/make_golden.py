"""Generate tests/golden/*.npz from the REAL reference (run in the build container only).

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden micro tiny # selected cases

For every case it (1) builds the seeded synthetic model (oracle/synth.py), (2) runs the vendored
reference itself — ``WhisperForConditionalGeneration.generate`` (modeling_whisper.py:1555-1568 ->
generation/utils.py:1185-1529), its encoder and its teacher-forced forward — and (3) runs the
restatement (oracle/whisper_ref.py) on the same tensors and records how far apart they are.
The .npz holds sub-sampled reference outputs, small enough to commit.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

from . import hf_reference as HF
from . import synth
from . import whisper_ref as R

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (size, batch, weight seed, mel seed, max_length, logit steps kept)
CASES = {
    "micro": ("micro", 3, 0, 1234, 448, [0, 1, 2, 3, 7, 31, 100, 446]),
    "tiny": ("tiny.en", 2, 0, 1234, 448, [0, 1, 2, 3, 7, 31, 100, 446]),
    "tiny_b1": ("tiny.en", 1, 0, 77, 448, [0, 1, 2, 5, 446]),          # BASELINE.json configs[0]
    "base_b16": ("base.en", 16, 0, 6, 448, [0, 1, 2, 9, 200, 446]),  # BASELINE.json configs[1]
    "small": ("small.en", 2, 0, 1234, 33, [0, 1, 2, 9, 31]),
    "medium": ("medium.en", 2, 0, 1234, 25, [0, 1, 2, 9, 23]),
    # the BENCHMARKED large-batch regime: B * H > 2 * SMs (warp-per-item paged self-attention, multi-kernel /
    # fused-chain decode step at d = 1024) and BASELINE.json configs[2] (small.en, batch 64)
    "medium_b24": ("medium.en", 24, 0, 4321, 25, [0, 1, 2, 9, 23]),
    "small_b64": ("small.en", 64, 0, 99, 17, [0, 1, 2, 9, 15]),
}
ENC_T_STRIDE, ENC_D_STRIDE, LOGIT_STRIDE = 97, 7, 53


def subsample_enc(enc: torch.Tensor) -> torch.Tensor:
    return enc[:, ::ENC_T_STRIDE, ::ENC_D_STRIDE].contiguous()


def subsample_logits(lg: torch.Tensor) -> torch.Tensor:
    return lg[:, ::LOGIT_STRIDE].contiguous()


def run_case(name: str) -> dict:
    size, B, wseed, mseed, max_length, steps_kept = CASES[name]
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = synth.make_config(size, max_length=max_length)
    sd = synth.make_weights(cfg, seed=wseed)
    mel = synth.make_mel(B, seed=mseed)
    model = HF.build_reference_model(cfg, sd)

    t0 = time.time()
    ref_ids = HF.reference_generate(model, mel, max_length=max_length)
    t_ref = time.time() - t0
    ref_enc = HF.reference_encode(model, mel)
    n_steps = ref_ids.shape[1] - 1
    ref_logits, ref_past = HF.reference_step_logits(model, ref_enc, ref_ids, n_steps)

    t0 = time.time()
    my_ids, my_enc, my_logits = R.greedy(mel, sd, cfg, return_logits=True)
    t_mine = time.time() - t0

    tokens_equal = bool(torch.equal(ref_ids, my_ids))
    enc_diff = float((ref_enc - my_enc).abs().max())
    logit_diff = max(float((a - b).abs().max()) for a, b in zip(ref_logits, my_logits))
    # decision margin of the reference at every free step (after the logits processors)
    gaps = []
    for n, lg in enumerate(ref_logits):
        top2 = R.process_logits(lg, n + 1, cfg).topk(2, dim=-1).values
        g = top2[:, 0] - top2[:, 1]
        gaps.append(float(g[torch.isfinite(g)].min()) if torch.isfinite(g).any() else float("inf"))
    min_gap = min(gaps)
    # cross K/V of the reference (layer 0 and last), sub-sampled, pins the cross-KV projection
    ck0, cv0 = ref_past[0][2], ref_past[0][3]
    ckL, cvL = ref_past[-1][2], ref_past[-1][3]
    skL = ref_past[-1][0]

    kept = [s for s in steps_kept if s < n_steps]
    arrays = dict(
        tokens=ref_ids.numpy().astype(np.int32),
        enc_sub=subsample_enc(ref_enc).numpy(),
        logits_sub=np.stack([subsample_logits(ref_logits[s]).numpy() for s in kept]),
        logit_steps=np.array(kept, dtype=np.int32),
        cross_k0_sub=ck0[:, :, ::ENC_T_STRIDE, ::ENC_D_STRIDE].contiguous().numpy(),
        cross_vL_sub=cvL[:, :, ::ENC_T_STRIDE, ::ENC_D_STRIDE].contiguous().numpy(),
        self_kL_sub=skL[:, :, ::11, ::ENC_D_STRIDE].contiguous().numpy(),
        top1_gap=np.array(gaps, dtype=np.float32),
    )
    meta = dict(
        case=name, size=size, batch=B, weight_seed=wseed, mel_seed=mseed, max_length=max_length,
        weights_fingerprint=synth.weights_fingerprint(sd),
        restatement_tokens_equal=tokens_equal, restatement_enc_maxabs=enc_diff,
        restatement_logit_maxabs=logit_diff, min_top1_gap=min_gap,
        distinct_tokens_per_row=[len(set(r.tolist())) for r in ref_ids],
        reference_generate_seconds=round(t_ref, 2), restatement_seconds=round(t_mine, 2),
        torch=torch.__version__, host_threads=torch.get_num_threads(),
        enc_stride=[ENC_T_STRIDE, ENC_D_STRIDE], logit_stride=LOGIT_STRIDE,
        reference="simplified HF Whisper, /root/reference/transformers/src/transformers/models/whisper/modeling_whisper.py",
    )
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **arrays)
    with open(os.path.join(GOLDEN_DIR, f"{name}.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return meta


def main(argv):
    names = argv or list(CASES)
    for n in names:
        m = run_case(n)
        print(json.dumps({k: m[k] for k in ("case", "restatement_tokens_equal", "restatement_enc_maxabs",
                                              "restatement_logit_maxabs", "min_top1_gap",
                                              "distinct_tokens_per_row", "reference_generate_seconds")}))
        assert m["restatement_tokens_equal"], "restatement diverged from the reference"


if __name__ == "__main__":
    main(sys.argv[1:])
