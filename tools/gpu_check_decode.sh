#!/bin/bash
# Dev tool (GPU box): decode parity (both variants), chain and mapper tests, then CUPTI timelines and the mapping step.
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 150 --timeout-method thread"
timeout 600 $PT tests/test_decode_gpu.py tests/test_reference_chain_gpu.py tests/test_mapper_gpu.py > gpurun_out/dec_all.log 2>&1; echo "decode+chain+mapper rc=$?"
tail -n 3 gpurun_out/dec_all.log
for v in 1 2; do
  SEGS_DECODE_VARIANT=$v timeout 200 python tools/timeline_decode.py > gpurun_out/tl_decode_C3_v$v.log 2>&1
  echo "== variant $v, C3"; head -n 7 gpurun_out/tl_decode_C3_v$v.log | tail -n 4
done
SEGS_DECODE_VARIANT=2 timeout 200 python tools/timeline_mapping.py 8 fused > gpurun_out/tl_mapping_v2.log 2>&1
grep -E "decode_|span" gpurun_out/tl_mapping_v2.log | head -8
bash tools/gpu_map_ab.sh | grep -E "==|value" | sed -E 's/.*"lanes": ([0-9]+).*"value": ([0-9.]+).*/lanes \1: \2 keyframes\/s/'
