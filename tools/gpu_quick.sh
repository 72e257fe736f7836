#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 150 --timeout-method thread"
timeout 600 $PT tests/test_decode_gpu.py tests/test_reference_chain_gpu.py > gpurun_out/dec_all.log 2>&1; echo "decode+chain rc=$?"; tail -n 1 gpurun_out/dec_all.log
SEGS_DECODE_VARIANT=2 timeout 200 python tools/timeline_decode.py 2>&1 | grep -E "span|decode_"
SEGS_DECODE_VARIANT=2 timeout 200 python tools/bench_mapping.py --steps 10 --lanes 4 2>&1 | tail -n 1 | sed -E 's/.*"lanes": ([0-9]+).*"value": ([0-9.]+).*/lanes \1: \2 keyframes\/s/'
SEGS_DECODE_VARIANT=2 timeout 200 python tools/bench_mapping.py --steps 10 --lanes 1 2>&1 | tail -n 1 | sed -E 's/.*"lanes": ([0-9]+).*"value": ([0-9.]+).*/lanes \1: \2 keyframes\/s/'
