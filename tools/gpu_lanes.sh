#!/bin/bash
# Dev tool (GPU box): the C4 mapping step against the number of views in flight per GPU.
for l in 2 3 4 5 6 8; do
  timeout 200 python tools/bench_mapping.py --steps 10 --lanes $l 2>&1 | tail -n 1 | sed -E 's/.*"lanes": ([0-9]+).*"value": ([0-9.]+).*"ms_per_step": ([0-9.]+).*/lanes \1: \2 keyframes\/s (\3 ms per 64-view step)/'
done | tee gpurun_out/lanes.log
