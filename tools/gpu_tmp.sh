#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "refill or session_options or finished_rows" > gpurun_out/r2v_pytest.log 2>&1
echo "rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2v_pytest.log | head -30
