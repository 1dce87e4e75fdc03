#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
python tools/chain_debug.py tiny.en 20
python tools/chain_debug.py tiny.en 127
python tools/chain_debug.py medium.en 24
} > gpurun_out/r2d_debug.log 2>&1
grep -v Warning gpurun_out/r2d_debug.log | tail
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused_chain or large_batch_paths or partial_batch" > gpurun_out/r2d_pytest_chain.log 2>&1
echo "pytest chain rc=$?" >> gpurun_out/r2d_pytest_chain.log
grep -E "^(FAILED|PASSED|ERROR)|passed|failed|^E  " gpurun_out/r2d_pytest_chain.log | head -40
for chain in 1; do
  timeout 600 python tools/decode_step_bench.py --batches 32,64,128,256 --lengths 128 --chain $chain > gpurun_out/r2d_step_chain$chain.md 2> gpurun_out/r2d_step_chain$chain.err
  cat gpurun_out/r2d_step_chain$chain.md; tail -5 gpurun_out/r2d_step_chain$chain.err
done
