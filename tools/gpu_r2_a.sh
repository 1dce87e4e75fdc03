#!/bin/bash
# round 2, GPU call A: new parity tests of the benchmarked large-batch path + fused chains, then the step microbench A/B
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(fused_chain or large_batch_paths or partial_batch) and not small_b64" > gpurun_out/r2a_pytest_chain.log 2>&1
echo "pytest chain rc=$?" >> gpurun_out/r2a_pytest_chain.log
tail -30 gpurun_out/r2a_pytest_chain.log
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "paged_self or cross_attention_large" > gpurun_out/r2a_pytest_ops.log 2>&1
echo "pytest ops rc=$?" >> gpurun_out/r2a_pytest_ops.log
tail -15 gpurun_out/r2a_pytest_ops.log
for chain in 0 1; do
  timeout 600 python tools/decode_step_bench.py --batches 32,64,128,256 --lengths 128,436 --chain $chain > gpurun_out/r2a_step_chain$chain.md 2> gpurun_out/r2a_step_chain$chain.err
  echo "step bench chain=$chain rc=$?"
  cat gpurun_out/r2a_step_chain$chain.md; tail -5 gpurun_out/r2a_step_chain$chain.err
done
