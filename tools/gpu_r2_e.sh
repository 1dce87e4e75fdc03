#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/chain_trace.py --batch 32 > gpurun_out/r2e_trace_b32.md 2> gpurun_out/r2e_trace_b32.err; tail -3 gpurun_out/r2e_trace_b32.err; cat gpurun_out/r2e_trace_b32.md
python tools/chain_trace.py --batch 256 > gpurun_out/r2e_trace_b256.md 2> gpurun_out/r2e_trace_b256.err; tail -3 gpurun_out/r2e_trace_b256.err; cat gpurun_out/r2e_trace_b256.md
