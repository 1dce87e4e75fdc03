#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/chain_debug.py medium.en 24 2>&1 | grep -v Warn | tail -2
python tools/chain_debug.py tiny.en 33 2>&1 | grep -v Warn | tail -2
python tools/chain_trace.py --batch 32 > gpurun_out/r2g_trace_b32.md 2> gpurun_out/r2g_trace_b32.err; tail -3 gpurun_out/r2g_trace_b32.err; cat gpurun_out/r2g_trace_b32.md
timeout 600 python tools/decode_step_bench.py --batches 32,64,128 --lengths 128 --chain 1 > gpurun_out/r2g_step_chain1.md 2> gpurun_out/r2g_step_chain1.err
cat gpurun_out/r2g_step_chain1.md; tail -5 gpurun_out/r2g_step_chain1.err
WB_BENCH_DEV=1 timeout 900 python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-probe --no-microbench --breakdown > gpurun_out/r2g_bench_b32.json 2> gpurun_out/r2g_bench_b32.err
grep breakdown gpurun_out/r2g_bench_b32.err; cat gpurun_out/r2g_bench_b32.json | cut -c1-600
