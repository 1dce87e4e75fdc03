#!/bin/bash
# Full GPU check of the working tree: every -m gpu test, smoke(), the default bench line (N = 1) and the CPU reference arm.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-suite}
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest_gpu.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -4 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
