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
    trace = torch.zeros(8 * (n_phases + 1), dtype=torch.int64, device=dev)
    _abi.call("wb_set_small_batch_path", 1)
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
    total_cycles = t[8 * (n_phases - 1) + 4] - t[0]
    mhz = float(os.popen("nvidia-smi --query-gpu=clocks.sm --format=csv,noheader,nounits -i 0").read().split()[0] or 0)
    print(f"# Whole-step kernel phase trace: {a.size} bf16, batch {a.batch}, length {a.length + 8}\n")
    print(f"step (CUDA events, graph replay) {step_us:.0f} us; kernel {total_cycles} SM cycles; nvidia-smi SM clock after the run {mhz:.0f} MHz\n")
    print("Mean SM cycles of CTA 0 per phase kind (linear layers: stage activations (LayerNorm / copy) -> block barrier -> request first "
          "weight tile -> weight stream + MMA + reduction + epilogue; then arrive + prefetch of the next phase -> wait for the grid).\n")
    print("| phase | count | stage | block barrier | first tile issue | stream + epilogue | work total | arrive + prefetch | grid wait | share of kernel |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    acc = [[0] * 7 for _ in range(9)]
    count = [0] * 9
    for ph in range(n_phases):
        k = 8 if ph == n_phases - 1 else ph % 8
        s0 = t[8 * ph: 8 * ph + 9]
        linear = k not in (1, 4)
        if linear:
            acc[k][0] += s0[1] - s0[0]
            acc[k][1] += s0[2] - s0[1]
            acc[k][2] += s0[3] - s0[2]
            acc[k][3] += s0[4] - s0[3]
        acc[k][4] += s0[4] - s0[0]
        if ph + 1 < n_phases:
            acc[k][5] += s0[5] - s0[4]
            acc[k][6] += s0[8] - s0[5]
        count[k] += 1
    for k in range(9):
        c = count[k]
        share = (acc[k][4] + acc[k][5] + acc[k][6]) / max(total_cycles, 1)
        cells = " | ".join(f"{acc[k][i] / c:.0f}" for i in range(7))
        print(f"| {KINDS[k]} | {c} | {cells} | {share:.1%} |")
    _abi.call("wb_set_step_trace", None)
    eng.close()


if __name__ == "__main__":
    main()
