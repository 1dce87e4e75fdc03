#!/bin/bash
# Round-2 evidence in ONE GPU call: same-box A/B of the decode-step paths, ncu launch lists of one decode step (B = 256 and 32)
# and one encoder chunk, one ncu --set full capture of the decode chains and of the decode attention kernels.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for chain in 0 1; do
  timeout 900 python tools/decode_step_bench.py --batches 17,32,64,128,256 --lengths 1,128,256,436 --chain $chain > gpurun_out/r02_step_chain$chain.md 2> gpurun_out/r02_step_chain$chain.err
  echo "step bench chain=$chain rc=$?"; grep "^| [0-9]" gpurun_out/r02_step_chain$chain.md
done
python tools/chain_trace.py --batch 32 > gpurun_out/r02_chain_trace_b32_final.md 2>/dev/null; python tools/chain_trace.py --batch 256 > gpurun_out/r02_chain_trace_b256_final.md 2>/dev/null
python tools/profile_step.py --batch 256 --region both --steps 1 > gpurun_out/r02_plain_b256.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b256.csv python tools/profile_step.py --batch 256 --region both --steps 1 > gpurun_out/r02_ncu_l256.log 2>&1
echo "launch list b256 rc=$?"
python tools/profile_step.py --batch 32 --region decode --steps 1 > gpurun_out/r02_plain_b32.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b32.csv python tools/profile_step.py --batch 32 --region decode --steps 1 > gpurun_out/r02_ncu_l32.log 2>&1
echo "launch list b32 rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"decode_chain|decode_attn|self_attn" -c 8 -f -o gpurun_out/r02_decode_b256 python tools/profile_step.py --batch 256 --region decode --steps 1 > gpurun_out/r02_ncu_full_b256.log 2>&1
echo "ncu full b256 rc=$?"
ncu -i gpurun_out/r02_decode_b256.ncu-rep --page raw --csv > gpurun_out/r02_decode_b256_raw.csv 2>/dev/null
ls -la gpurun_out | grep r02_
