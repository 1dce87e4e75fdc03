#!/usr/bin/env python
"""Where the whole-step decoder kernel (csrc/step_mega.cu) spends its time: CTA 0 stamps its SM clock after every phase and
after every grid barrier (wb_set_step_trace); this prints the mean per phase kind for one decode step at length t.

    python tools/step_trace.py [--size medium.en] [--batch 1] [--length 128] > profiles/r01_step_trace_b1.md
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KINDS = ["LN1+qkv", "self-attention", "out-proj", "LN2+cross-q", "cross-attention", "cross-out", "LN3+fc1", "fc2", "LN+LM head"]


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--size", default="medium.en")
    p.add_argument("--batch", type=int, default=1)
    p.add_argument("--length", type=int, default=128)
    a = p.parse_args()
    from whisper_trtllm_b200 import WhisperEngine, _abi, synthetic as synth
    from whisper_trtllm_b200._abi import ptr

    dev = torch.device("cuda", 0)
    cfg = synth.make_config(a.size)
    sd = synth.make_weights(cfg, seed=0)
    L = cfg["decoder_layers"]
    n_phases = 8 * L + 1
    trace = torch.zeros(2 * n_phases + 1, dtype=torch.int64, device=dev)
    _abi.call("wb_set_small_batch_path", 2)
    _abi.call("wb_set_step_trace", ptr(trace))
    eng = WhisperEngine(cfg, sd, dtype="bf16", max_batch=a.batch, enc_chunk=min(a.batch, 32), device=dev)
    del sd
    eng.encode(synth.make_mel(a.batch, seed=1234).to(dev), return_hidden=False)
    eng.decode_begin(a.batch)
    eng.decode_run(max_steps=a.length, check_every=1 << 20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.decode_run(max_steps=8, check_every=1 << 20)
    e1.record()
    torch.cuda.synchronize()
    step_us = e0.elapsed_time(e1) / 8 * 1e3
    t = trace.cpu().tolist()
    total_cycles = t[2 * n_phases - 1] - t[0]
    # the kernel is all but the argmax kernel and the launch gap of a step: calibrate the SM clock on a second measurement
    mhz = float(os.popen("nvidia-smi --query-gpu=clocks.sm --format=csv,noheader,nounits -i 0").read().split()[0] or 0)
    print(f"# Whole-step kernel phase trace: {a.size} bf16, batch {a.batch}, length {a.length + 8}\n")
    print(f"step (CUDA events, graph replay) {step_us:.0f} us; kernel {total_cycles} SM cycles; nvidia-smi SM clock after the run {mhz:.0f} MHz\n")
    print("| phase | count | work cycles (mean) | barrier cycles (mean) | share of kernel |")
    print("|---|---|---|---|---|")
    work = [0] * 9
    barrier = [0] * 9
    count = [0] * 9
    for ph in range(n_phases):
        k = 8 if ph == n_phases - 1 else ph % 8
        start = t[2 * ph]
        work[k] += t[2 * ph + 1] - start
        if ph + 1 < n_phases:
            barrier[k] += t[2 * ph + 2] - t[2 * ph + 1]
        count[k] += 1
    for k in range(9):
        share = (work[k] + barrier[k]) / max(total_cycles, 1)
        print(f"| {KINDS[k]} | {count[k]} | {work[k] / count[k]:.0f} | {barrier[k] / count[k]:.0f} | {share:.1%} |")
    print(f"\nwork = phase body of CTA 0 (staging, weight stream, MMA, epilogue); barrier = prefetch of the next phase + waiting for the slowest CTA + release / acquire round trips.")
    _abi.call("wb_set_step_trace", None)
    eng.close()


if __name__ == "__main__":
    main()
