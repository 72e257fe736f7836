#!/bin/bash
# Dev tool (GPU box): the 64-keyframe mapping step with both decode variants, 4 and 1 lanes.
mkdir -p gpurun_out
for v in 1 2; do for l in 4 1; do
  echo "== variant $v lanes $l"; SEGS_DECODE_VARIANT=$v timeout 200 python tools/bench_mapping.py --steps 10 --lanes $l 2>&1 | tail -n 1
done; done | tee gpurun_out/map_ab.log
