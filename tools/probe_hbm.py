#!/usr/bin/env python
"""Pure-read HBM ceiling of this GPU (tools, not product): streams a 16 GiB buffer with the library's probe kernels and
prints GB/s per mode / CTAs-per-SM, next to MEASURED_PEAKS.json's copy figure."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whisper_trtllm_b200 import _abi  # noqa: E402
from whisper_trtllm_b200._abi import c_size_t, ptr, stream_handle  # noqa: E402

dev = torch.device("cuda", 0)
nbytes = 16 << 30
buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
buf.random_(0, 255)
sink = torch.zeros(4, dtype=torch.uint8, device=dev)
res = {}
for mode, cps_list in ((0, (2, 4, 8)), (1, (1, 2))):
    for cps in cps_list:
        best = 0.0
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _abi.call("wb_bandwidth_probe", ptr(buf), c_size_t(nbytes), mode, cps, ptr(sink), stream_handle())
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        res[f"mode{mode}_ctas{cps}"] = round(best, 1)
        print(f"mode {mode} ({'ldg.128' if mode == 0 else 'cp.async.bulk 16KB'}), {cps} CTAs/SM: {best:.1f} GB/s", flush=True)
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(peaks):
    print("MEASURED_PEAKS hbm_gbs (copy):", json.load(open(peaks))["hbm_gbs"])
print(json.dumps(res))
