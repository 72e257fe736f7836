#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 300 $NCU -k regex:decode_wgrad_tc -s 2 -c 1 -o gpurun_out/ncu_wg python tools/bench_decode.py 200000 --no-oracle > gpurun_out/ncu_wg.log 2>&1; echo "rc=$?"
