#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
python tools/chain_debug.py tiny.en 20
WB_CHAIN_SERIAL=1 python tools/chain_debug.py tiny.en 20
WB_CHAIN_BN=64 python tools/chain_debug.py tiny.en 20
WB_CHAIN_BN=128 python tools/chain_debug.py tiny.en 20
WB_CHAIN_SPLITS=1 python tools/chain_debug.py tiny.en 20
WB_CHAIN_BN=64 WB_CHAIN_SPLITS=1 python tools/chain_debug.py tiny.en 20
WB_CHAIN_BN=32 python tools/chain_debug.py tiny.en 130
WB_CHAIN_BN=64 WB_CHAIN_SPLITS=1 python tools/chain_debug.py tiny.en 128
WB_CHAIN_BN=64 WB_CHAIN_SPLITS=1 python tools/chain_debug.py tiny.en 127
} > gpurun_out/r2c_debug.log 2>&1
grep -v Warning gpurun_out/r2c_debug.log | tail -40
