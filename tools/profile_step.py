#!/usr/bin/env python
"""Profiling harness for ncu (B200_PROFILING.md recipe): sets up the bench workload, warms it up, then runs
exactly the region of interest between cudaProfilerStart/Stop so that

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
    ncu --profile-from-start off --set full -k regex:<kernel> -c 3 ...

see only those launches.  Regions:
  decode   N decode steps at sequence length ~`--at-len` (default: 2 steps at length 224 = the mean length of the
           447-step loop, so the self-attention share is representative)
  encode   one encoder chunk (stem + layers + cross-K/V projection) of `--enc-chunk` utterances
Without ncu the script just runs and prints CUDA-event timings of the same regions (the plain run that must
exit 0 before any ncu run).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--size", default="medium.en")
    p.add_argument("--batch", type=int, default=256)
    p.add_argument("--dtype", default="bf16")
    p.add_argument("--enc-chunk", type=int, default=32)
    p.add_argument("--region", default="decode", choices=["decode", "encode", "both"])
    p.add_argument("--at-len", type=int, default=224)
    p.add_argument("--steps", type=int, default=2)
    a = p.parse_args()

    from whisper_trtllm_b200 import WhisperEngine, synthetic as synth

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = synth.make_config(a.size)
    sd = synth.make_weights(cfg, seed=0)
    B = a.batch
    eng = WhisperEngine(cfg, sd, dtype=a.dtype, max_batch=B, enc_chunk=min(a.enc_chunk, B), device=dev)
    del sd
    mel = synth.make_mel(B, seed=1234).to(dev)
    chunk = mel[:min(a.enc_chunk, B)].contiguous()

    # warm-up: full encoder once (fills the cross K/V of every row), decode up to the length of interest
    eng.encode(mel, return_hidden=False)
    eng.decode_begin(B)
    eng.decode_run(max_steps=a.at_len - 1, check_every=1 << 20)
    torch.cuda.synchronize()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.profiler.start()
    if a.region in ("decode", "both"):
        ev[0].record()
        for _ in range(a.steps):
            eng.decode_step()
        ev[1].record()
    if a.region in ("encode", "both"):
        ev[2].record()
        eng.encode(chunk, return_hidden=False)
        ev[3].record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    if a.region in ("decode", "both"):
        print(f"decode: {a.steps} steps at len {a.at_len}: {ev[0].elapsed_time(ev[1]) / a.steps:.3f} ms/step (B={B}, {a.size}, {a.dtype})")
    if a.region in ("encode", "both"):
        print(f"encode: chunk of {chunk.shape[0]}: {ev[2].elapsed_time(ev[3]):.3f} ms")
    print(f"launches so far: {eng.launch_count()}")
    eng.close()


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"done in {time.time() - t0:.1f}s")
