#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/chain_debug.py medium.en 24 2>&1 | grep -v Warn | tail -1
python tools/chain_debug.py tiny.en 33 2>&1 | grep -v Warn | tail -1
python tools/chain_trace.py --batch 32 > gpurun_out/r2h_trace_b32.md 2> gpurun_out/r2h_trace_b32.err; tail -3 gpurun_out/r2h_trace_b32.err; cat gpurun_out/r2h_trace_b32.md
for items in 296 1024 2048 4096; do
  echo "== WB_SELF_CTA_ITEMS=$items"
  WB_SELF_CTA_ITEMS=$items timeout 600 python tools/decode_step_bench.py --batches 32,64,128,256 --lengths 128,436 --chain 1 2> gpurun_out/r2h_step.err | grep "^| [0-9]"
done
echo "== chain 0 (WB_SELF_CTA_ITEMS=1024)"
timeout 600 python tools/decode_step_bench.py --batches 32,64,128,256 --lengths 128,436 --chain 0 2> gpurun_out/r2h_step.err | grep "^| [0-9]"
