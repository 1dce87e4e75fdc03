#!/usr/bin/env python
"""BASELINE.json configs[4]: medium.en decoder-step microbench — device time of ONE greedy decode step (24 layers: cached
self-attention over t keys appended in place, cross-attention over 1500 frames, 144 skinny GEMMs, LM head, logits
processors + argmax) for batch 1..512 and t in {1, 64, 128, 256, 447}, CUDA-graph replay, CUDA events around 8 steps.

    python tools/decode_step_bench.py [--size medium.en] [--dtype bf16] > profiles/r01_decode_step_us.md
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--size", default="medium.en")
    p.add_argument("--dtype", default="bf16")
    p.add_argument("--batches", default="1,8,64,256,512")
    p.add_argument("--lengths", default="1,64,128,256,436")   # the 10 timed / warm steps must fit below max_length 448
    p.add_argument("--steps", type=int, default=8)
    p.add_argument("--small-path", type=int, default=-1, help="wb_set_small_batch_path: 1 = whole-step kernel for B <= 16 (library default), 2 = same with mma.sync attention, 0 = multi-kernel step")
    p.add_argument("--chain", type=int, default=-1, help="wb_set_decode_chain_path: 1 = fused GEMM / LayerNorm chains for B > 16 (library default), 0 = one kernel per GEMM / LayerNorm")
    a = p.parse_args()
    from whisper_trtllm_b200 import WhisperEngine, _abi, synthetic as synth
    if a.small_path >= 0:
        _abi.call("wb_set_small_batch_path", a.small_path)
    if a.chain >= 0:
        _abi.call("wb_set_decode_chain_path", a.chain)

    dev = torch.device("cuda", 0)
    cfg = synth.make_config(a.size)
    sd = synth.make_weights(cfg, seed=0)
    batches = [int(x) for x in a.batches.split(",")]
    lengths = [int(x) for x in a.lengths.split(",")]
    Bmax = max(batches)
    eng = WhisperEngine(cfg, sd, dtype=a.dtype, max_batch=Bmax, enc_chunk=32, device=dev)
    del sd
    mel = synth.make_mel(min(Bmax, 64), seed=1234).to(dev)
    # every row's cross K/V holds a real projection (the same 64 utterances repeated: timing does not depend on the values)
    big = mel.repeat((Bmax + mel.shape[0] - 1) // mel.shape[0], 1, 1)[:Bmax].contiguous()
    eng.encode(big, return_hidden=False)
    torch.cuda.synchronize()
    H, d, L = cfg["decoder_attention_heads"], cfg["d_model"], cfg["decoder_layers"]
    es = 2 if a.dtype == "bf16" else 4
    print(f"# Decode-step time, {a.size} {a.dtype}, one B200 (us per step, CUDA events over {a.steps} graph-replayed steps)\n")
    print("HBM floor = (decoder weights 0.81 GB + B x (cross K/V 147.5 MB + self K/V 2 x 24 x t x 1024 x 2 B)) / 6.54 TB/s.\n")
    print("| batch | " + " | ".join(f"t = {t}" for t in lengths) + " |")
    print("|---|" + "---|" * len(lengths))
    for B in batches:
        cells = []
        for t in lengths:
            eng.decode_begin(B)
            if t > 1:
                eng.decode_run(max_steps=t - 1, check_every=1 << 20)
            eng.decode_run(max_steps=2, check_every=1 << 20)       # warm (graph for this batch size)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            eng.decode_run(max_steps=a.steps, check_every=1 << 20)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / a.steps * 1e3
            tt = t + 2 + a.steps / 2
            floor = (405.8e6 * es + B * (2 * L * 1500 * d * es + 2 * L * tt * d * es)) / 6538.6e9 * 1e6
            cells.append(f"{us:.0f} (floor {floor:.0f})")
        print(f"| {B} | " + " | ".join(cells) + " |", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
