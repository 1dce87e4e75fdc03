#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/chain_debug.py medium.en 24 2>&1 | grep -v Warn | tail -1
python tools/chain_debug.py tiny.en 33 2>&1 | grep -v Warn | tail -1
python tools/chain_debug.py base.en 130 2>&1 | grep -v Warn | tail -1
python tools/chain_trace.py --batch 32 > gpurun_out/r2k_trace_b32.md 2> gpurun_out/r2k_trace_b32.err; tail -3 gpurun_out/r2k_trace_b32.err; cat gpurun_out/r2k_trace_b32.md
python tools/chain_trace.py --batch 256 > gpurun_out/r2k_trace_b256.md 2> gpurun_out/r2k_trace_b256.err; tail -3 gpurun_out/r2k_trace_b256.err; head -20 gpurun_out/r2k_trace_b256.md
timeout 600 python tools/decode_step_bench.py --batches 32,64,128,256 --lengths 128,436 --chain 1 2> gpurun_out/r2k_step.err | grep "^| [0-9]"
