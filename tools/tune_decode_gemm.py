#!/usr/bin/env python
"""Empirical (BLOCK_N, split-K) sweep for the skinny decode-step GEMMs (tools, not product): times wb_linear_splitk for the
four decoder shapes at M = 256 with 64 distinct weight matrices per shape (so the weights come from HBM, like in the real
step) and prints us per launch per configuration, next to what pick_config chooses on its own."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whisper_trtllm_b200 import _abi  # noqa: E402
from whisper_trtllm_b200._abi import ptr, stream_handle  # noqa: E402

dev = torch.device("cuda", 0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SHAPES = [("attn/out/cq N1024 K1024 (split ok)", 1024, 1024, True), ("fc2 N1024 K4096 (split ok)", 1024, 4096, True),
          ("qkv N3072 K1024", 3072, 1024, False), ("fc1 N4096 K1024", 4096, 1024, False)]
COPIES, ITERS = 64, 256


def run(N, K, bn, s, Ws, A, parts):
    _abi.call("wb_set_gemm_block_n", bn)
    chosen = ctypes.c_int(0)
    def once(i):
        _abi.call("wb_linear_splitk", ptr(A), K, ptr(Ws[i % COPIES]), K, _abi.BF16, ptr(parts), M * N, M, N, K, s, 8,
                  ctypes.byref(chosen), stream_handle())
    for i in range(16):
        once(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(ITERS):
        once(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / ITERS * 1e3, chosen.value


for name, N, K, split_ok in SHAPES:
    Ws = [(torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16) for _ in range(COPIES)]
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    parts = torch.empty(8, M, N, device=dev)
    auto_us, auto_s = run(N, K, 0, 0 if split_ok else 1, Ws, A, parts)
    print(f"== {name}, M={M}: auto choice {auto_us:.2f} us (splits {auto_s})", flush=True)
    for bn in (32, 64, 128, 256):
        row = []
        for s in ((1, 2, 4, 8) if split_ok else (1,)):
            if (K // 64) % s:
                continue
            us, _ = run(N, K, bn, s, Ws, A, parts)
            row.append(f"S={s}: {us:6.2f}")
        print(f"   BN={bn:3d}  " + "  ".join(row), flush=True)
    del Ws
_abi.call("wb_set_gemm_block_n", 0)
