#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/r2l_pytest_ops.log 2>&1
echo "ops rc=$?"; tail -5 gpurun_out/r2l_pytest_ops.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r2l_pytest_parity.log 2>&1
echo "parity rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2l_pytest_parity.log | head -20
WB_BENCH_DEV=1 timeout 900 python bench.py --batch 256 --steps 2 --warmup 2 --no-cpu-baseline --no-microbench --breakdown --breakdown-only enc_gemm,enc_attn,cross_kv,dec_gemm,self_attn,cross_attn > gpurun_out/r2l_bench_b256.json 2> gpurun_out/r2l_bench_b256.err
echo "bench rc=$?"; grep -E "breakdown|probe|warmup" gpurun_out/r2l_bench_b256.err; cat gpurun_out/r2l_bench_b256.json
