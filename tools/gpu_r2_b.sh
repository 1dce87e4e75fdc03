#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused_chain or large_batch_paths or partial_batch" > gpurun_out/r2b_pytest_chain.log 2>&1
echo "pytest chain rc=$?" >> gpurun_out/r2b_pytest_chain.log
grep -E "^(FAILED|PASSED|ERROR)|passed|failed|^E  " gpurun_out/r2b_pytest_chain.log | head -60
