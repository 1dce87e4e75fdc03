#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/chain_debug.py medium.en 24 2>&1 | grep -v Warn | tail -2
python tools/chain_trace.py --batch 32 > gpurun_out/r2f_trace_b32.md 2> gpurun_out/r2f_trace_b32.err; tail -3 gpurun_out/r2f_trace_b32.err; cat gpurun_out/r2f_trace_b32.md
python tools/chain_trace.py --batch 256 > gpurun_out/r2f_trace_b256.md 2> gpurun_out/r2f_trace_b256.err; tail -3 gpurun_out/r2f_trace_b256.err; cat gpurun_out/r2f_trace_b256.md
timeout 600 python tools/decode_step_bench.py --batches 32,64,128,256 --lengths 128 --chain 1 > gpurun_out/r2f_step_chain1.md 2> gpurun_out/r2f_step_chain1.err
cat gpurun_out/r2f_step_chain1.md; tail -5 gpurun_out/r2f_step_chain1.err
