#!/usr/bin/env python
"""Smallest end-to-end exercise of every kernel, meant to run under `compute-sanitizer --tool memcheck` (one tool per
gpurun call, B200_PROFILING.md): micro model (d = 128, 2 layers, 2 heads), 3 utterances, a few greedy steps in fp32
(CUDA-core kernels) and bf16 (tcgen05 GEMM + flash attention, split-K, graph replay), the module-level drop-ins,
the bulk-ring attention kernel and the log-mel front-end.  Exits non-zero on any mismatch with the CPU oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/: it uses the oracle as the checker)
sys.path.insert(0, ROOT)
from oracle import logmel_ref as LM  # noqa: E402
from oracle import synth, whisper_ref as R  # noqa: E402
from whisper_trtllm_b200 import WhisperEngine, _abi  # noqa: E402
from whisper_trtllm_b200.frontend import LogMelFrontend  # noqa: E402
from whisper_trtllm_b200.model import decoder_from_config, load_decoder_from_hf  # noqa: E402


def main():
    dev = "cuda:0"
    steps = 6
    cfg = synth.make_config("micro", max_length=steps + 1)
    sd = synth.make_weights(cfg, seed=0)
    mel = synth.make_mel(3, seed=1234)
    ref = R.greedy(mel, sd, cfg)
    for dtype in ("float32", "bfloat16"):
        eng = WhisperEngine(cfg, sd, dtype=dtype, max_batch=3, enc_chunk=2, device=dev)
        ids = eng.generate(mel.to(dev)).cpu().long()
        ids2 = eng.generate(mel.to(dev)).cpu().long()          # second run replays the captured graph
        torch.cuda.synchronize()
        assert torch.equal(ids, ids2)
        if dtype == "float32":
            assert torch.equal(ids, ref), (ids, ref)
        eng.close()
        print(f"{dtype}: engine ok ({eng.launch_count()} launches so far)", flush=True)
    # multi-stream (bulk-ring cross-attention + lean GEMM)
    eng = WhisperEngine(cfg, sd, dtype="bfloat16", max_batch=3, device=dev, n_streams=2)
    eng.generate(mel.to(dev))
    eng.close()
    _abi.call("wb_set_decode_attention_backend", 0)
    _abi.call("wb_set_lean_decode_gemm", 0)
    print("multi-stream ok", flush=True)
    # module-level drop-in: one decoder step with caches
    enc = R.encode(mel, sd, cfg).to(dev)
    dec = load_decoder_from_hf(decoder_from_config(cfg, dtype="bfloat16"), sd)
    L, H = cfg["decoder_layers"], cfg["decoder_attention_heads"]
    sk = torch.zeros(L, 3, H, 1, 64, device=dev); ck = torch.zeros(L, 3, H, 1500, 64, device=dev)
    logits, sk, sv, ck, cv = dec(ref[:, :1].to(dev), enc, sk, sk.clone(), ck, ck.clone(), torch.zeros(1, device=dev), torch.zeros(1, device=dev))
    logits, sk, sv, ck, cv = dec(ref[:, 1:2].to(dev), enc, sk, sv, ck, cv, torch.zeros(2, device=dev), torch.zeros(1501, device=dev))
    torch.cuda.synchronize()
    assert torch.isfinite(logits).all()
    print("module drop-in ok", flush=True)
    # log-mel front-end
    fe = LogMelFrontend(dev)
    wave = LM.synth_wave("chirp_short", seed=1)
    out = fe([wave]).cpu().numpy()[0]
    err = abs(out - LM.log_mel(wave)).max()
    assert err < 5e-3, err
    print(f"log-mel ok (max err {err:.2e})", flush=True)


if __name__ == "__main__":
    main()
    print("SANITIZE_SMALL_OK")
