#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/chain_debug.py medium.en 24 2>&1 | grep -v Warn | tail -1
for pdl in 0 1; do
  echo "== pdl $pdl"
  timeout 300 python tools/decode_step_bench.py --batches 32,64,256 --lengths 128,436 --chain 1 --pdl $pdl 2> gpurun_out/r2y_step_pdl$pdl.err | grep "^| [0-9]"
  tail -2 gpurun_out/r2y_step_pdl$pdl.err
done
