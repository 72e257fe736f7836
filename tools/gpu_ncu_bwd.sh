#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 300 $NCU -k regex:decode_backward_kernel -s 2 -c 1 -o gpurun_out/ncu_bwd_view python tools/one_view_mapping.py 2 > gpurun_out/ncu_bwd_view.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/ncu_bwd_view.log
