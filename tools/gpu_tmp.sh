#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" 2>&1 | tail -3
timeout 600 python tools/decode_step_bench.py --batches 32,64,128,256 --lengths 128,436 --chain 1 2> gpurun_out/r2u_step.err | grep "^| [0-9]"
WB_BENCH_DEV=1 timeout 900 python bench.py --batch 128 --steps 1 --warmup 1 --no-cpu-baseline --no-probe --no-microbench --breakdown --breakdown-only enc_attn,enc_gemm,cross_attn,self_attn,dec_gemm > gpurun_out/r2u_bench_b128.json 2> gpurun_out/r2u_bench_b128.err
grep breakdown gpurun_out/r2u_bench_b128.err; cut -c1-200 gpurun_out/r2u_bench_b128.json
