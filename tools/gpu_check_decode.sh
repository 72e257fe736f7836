#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 150 --timeout-method thread"
timeout 600 $PT tests/test_decode_gpu.py tests/test_reference_chain_gpu.py tests/test_mapper_gpu.py > gpurun_out/dec_all.log 2>&1; echo "decode+chain+mapper rc=$?"
SEGS_DECODE_VARIANT=2 timeout 200 python tools/timeline_decode.py > gpurun_out/tl_decode_C3_v2.log 2>&1
SEGS_DECODE_VARIANT=2 timeout 200 python tools/timeline_mapping.py 8 fused > gpurun_out/tl_mapping_v2.log 2>&1
tail -n 3 gpurun_out/dec_all.log
head -n 7 gpurun_out/tl_decode_C3_v2.log | tail -n 4
grep -E "decode_|span" gpurun_out/tl_mapping_v2.log | head -8
