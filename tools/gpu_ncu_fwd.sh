#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 300 $NCU -k regex:decode_forward_v2 -s 3 -c 1 -o gpurun_out/ncu_fwd_v2 python tools/bench_decode.py 200000 --no-oracle > gpurun_out/ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
tail -n 2 gpurun_out/ncu_fwd.log
