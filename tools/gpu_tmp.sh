#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for thr in 128 256; do
  echo "== WB_SELF_CTA_THREADS=$thr"
  WB_SELF_CTA_THREADS=$thr timeout 300 python tools/decode_step_bench.py --batches 32,64 --lengths 128,436 --chain 1 2> gpurun_out/r2z_step.err | grep "^| [0-9]"
done
