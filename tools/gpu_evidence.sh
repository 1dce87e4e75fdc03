#!/usr/bin/env bash
# One gpurun call that refreshes the whole evidence set of a build (run from the repo root ON the GPU box):
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh r2 v1'
#
# <round tag> <build tag> name the files written under gpurun_out/; summarise them afterwards in the build container with
# tools/summarize_launches.py and tools/summarize_ncu.py and copy the summaries to profiles/ (profiles/INDEX.md lists what goes
# where).  Order follows B200_PROFILING.md: every program first exits 0 WITHOUT ncu, a number printed under ncu is never used.
# Each step has its own timeout so that a hang cannot eat the call; a failing step does not stop the later ones.
set -u
R=${1:-r2}; V=${2:-v0}; O=gpurun_out; mkdir -p "$O"
# step <name> <seconds> <stdout file> <stderr file | -> (- = into the stdout file) <command...>
step() {
    local name=$1 limit=$2 out=$3 err=$4; shift 4
    echo "== $name"
    if [ "$err" = "-" ]; then timeout "$limit" "$@" > "$out" 2>&1; else timeout "$limit" "$@" > "$out" 2> "$err"; fi
    echo "== $name rc=$?"
}

step pytest 400       "$O/${R}_pytest_gpu_${V}.log" -  python -m pytest tests -x -q -m gpu
step smoke 120        "$O/${R}_smoke_${V}.log" -       python __graft_entry__.py smoke
step bench 400        "$O/${R}_bench_${V}.json" "$O/${R}_bench_${V}.err"  python bench.py --gpus 1 --steps 3 --warmup 3 --breakdown
step decode_steps 300 "$O/${R}_decode_step_${V}.md" "$O/${R}_decode_step_${V}.err"  python tools/decode_step_bench.py
# plain run of the profiling harness, then the launch list, then ONE --set full capture per region
step prof_plain 200   "$O/${R}_prof_plain_${V}.log" -  python tools/profile_step.py --region both
step launches 400     "$O/${R}_launches_${V}.log" -    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file "$O/${R}_launches_${V}.csv" python tools/profile_step.py --region both --steps 1
step ncu_decode 500   "$O/${R}_ncu_decode_${V}.log" -  ncu --profile-from-start off --set full --clock-control none --import-source on -c 16 -f \
     -o "$O/${R}_dec_layer_${V}" python tools/profile_step.py --region decode --steps 1
step ncu_encode 500   "$O/${R}_ncu_encode_${V}.log" -  ncu --profile-from-start off --set full --clock-control none --import-source on -c 14 -f \
     -o "$O/${R}_enc_layer_${V}" python tools/profile_step.py --region encode
ls -la "$O" | tail -20
