#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "paged_cache or finished_rows or refill or whole_step_kernel_edge or bf16_teacher" 2>&1 | tail -3
timeout 600 python tools/decode_step_bench.py --batches 64,128,256 --lengths 128,256,436 --chain 1 2> gpurun_out/r3a_step.err | grep "^| [0-9]"
WB_BENCH_DEV=1 timeout 900 python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-probe --no-microbench --breakdown --breakdown-only self_attn,cross_attn > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err
grep breakdown gpurun_out/r3a_bench.err; cut -c1-140 gpurun_out/r3a_bench.json
