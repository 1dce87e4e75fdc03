#!/usr/bin/env python
"""Where the fused decode chains (csrc/step_chain.cu) spend their time: CTA 0 stamps its SM clock at 8 points of every phase
(wb_set_step_trace); prints the mean per phase kind for one decode step at length t.

    python tools/chain_trace.py [--size medium.en] [--batch 32] [--length 128]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NAMES = ["out-proj", "LN2", "cross-q", "cross-out", "LN3", "fc1", "fc2", "LN1'", "qkv'"]


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--size", default="medium.en")
    p.add_argument("--batch", type=int, default=32)
    p.add_argument("--length", type=int, default=128)
    a = p.parse_args()
    from whisper_trtllm_b200 import WhisperEngine, _abi, synthetic as synth
    from whisper_trtllm_b200._abi import ptr

    dev = torch.device("cuda", 0)
    cfg = synth.make_config(a.size)
    sd = synth.make_weights(cfg, seed=0)
    L = cfg["decoder_layers"]
    n_phases = 2 + 9 * L
    trace = torch.zeros(8192, dtype=torch.int64, device=dev)
    eng = WhisperEngine(cfg, sd, dtype="bf16", max_batch=a.batch, enc_chunk=min(a.batch, 32), device=dev)
    del sd
    eng.encode(synth.make_mel(a.batch, seed=1234).to(dev), return_hidden=False)
    eng.decode_begin(a.batch)
    eng.decode_run(max_steps=a.length, check_every=1 << 20)
    torch.cuda.synchronize()
    _abi.call("wb_set_step_trace", ptr(trace))
    _abi.call("wb_set_cuda_graphs", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.decode_run(max_steps=4, check_every=1 << 20)
    e1.record()
    torch.cuda.synchronize()
    _abi.call("wb_set_step_trace", None)
    _abi.call("wb_set_cuda_graphs", 1)
    t = trace.cpu().tolist()
    print(f"# Fused-chain phase trace: {a.size} bf16, batch {a.batch}, length {a.length + 4}; eager step {e0.elapsed_time(e1) / 4 * 1e3:.0f} us\n")
    print("SM cycles of CTA 0, mean over the layers.  GEMM phase: start -> [grid barrier seen by the epilogue leader] ; producer: barrier "
          "seen -> activation loads issued -> first k-block landed -> first accumulator complete -> share written -> arrived.\n")
    print("| phase | n | start->leader past barrier | start->producer past barrier | ->A issued | ->first k-block landed | ->accumulator | ->written | ->arrived | total |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    acc = {k: [0.0] * 8 for k in range(9)}
    cnt = {k: 0 for k in range(9)}
    for l in range(L):
        for k in range(9):
            ph = 2 + 9 * l + k
            if l == L - 1 and k == 8:
                continue
            s = t[8 * ph: 8 * ph + 8]
            if s[0] == 0:
                continue
            gemm = k not in (1, 4, 7)
            first = k in (0, 3)     # first phase of a launch: no barrier in front
            row = acc[k]
            row[0] += s[1] - s[0]
            if gemm:
                if not first:
                    row[1] += s[2] - s[0]
                    row[2] += s[3] - s[2]
                    row[3] += s[4] - s[3]
                else:
                    row[3] += s[4] - s[0]
                row[4] += s[5] - s[4]
                row[5] += s[6] - s[5]
            else:
                row[5] += s[6] - s[1]
            last = k in (2, 8) or (l == L - 1 and k == 7)
            if not last:
                row[6] += s[7] - s[6]
                row[7] += s[7] - s[0]
            else:
                row[7] += s[6] - s[0]
            cnt[k] += 1
    for k in range(9):
        c = max(cnt[k], 1)
        print(f"| {NAMES[k]} | {cnt[k]} | " + " | ".join(f"{v / c:.0f}" for v in acc[k]) + " |")
    print("\nLayerNorm phases, CTA 0 (cycles): leader past barrier -> loads issued -> landed + warp sum -> statistics done -> rows written")
    for k in (1, 4, 7):
        d = [0.0] * 4
        for l in range(L):
            ss = t[8 * (2 + 9 * l + k): 8 * (2 + 9 * l + k) + 8]
            d[0] += ss[2] - ss[1]; d[1] += ss[3] - ss[2]; d[2] += ss[4] - ss[3]; d[3] += ss[6] - ss[4]
        print(f"| {NAMES[k]} | " + " | ".join(f"{v / L:.0f}" for v in d) + " |")
    # per-launch spans: first stamp of the first phase to the last stamp of the last phase
    spans2 = [t[8 * (2 + 9 * l + 2) + 6] - t[8 * (2 + 9 * l)] for l in range(L)]
    spans3 = [t[8 * (2 + 9 * l + (8 if l < L - 1 else 7)) + 6] - t[8 * (2 + 9 * l + 3)] for l in range(L)]
    print(f"\nlaunch [out-proj, LN2, cross-q]: mean {sum(spans2) / L:.0f} cycles; launch [cross-out .. qkv']: mean {sum(spans3) / L:.0f} cycles "
          f"(from the epilogue leader's first stamp: the kernel prologue is not included)")
    if t[4096] != 0:
        base = t[8 * 16]
        print("\nring-stage timeline of phase 16 (fc1, layer 1; one stage = one group of k-blocks), cycles since the phase start of the "
              "epilogue leader:")
        print("| stage | data landed (MMA warp) | MMAs issued |")
        print("|---|---|---|")
        for kb in range(32):
            if t[4096 + 2 * kb] == 0:
                break
            print(f"| {kb} | {t[4096 + 2 * kb] - base} | {t[4096 + 2 * kb + 1] - base} |")
    eng.close()


if __name__ == "__main__":
    main()
