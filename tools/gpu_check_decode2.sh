#!/bin/bash
# Dev tool (GPU box): decode variants — parity, timelines, then ncu captures of the two new kernels.
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 150 --timeout-method thread"
timeout 600 $PT tests/test_decode_gpu.py tests/test_reference_chain_gpu.py tests/test_mapper_gpu.py > gpurun_out/dec_all.log 2>&1; echo "decode+chain+mapper rc=$?" | tee -a gpurun_out/dec_rc2.log
for v in 1 2; do
  SEGS_DECODE_VARIANT=$v timeout 200 python tools/timeline_decode.py > gpurun_out/tl_decode_C3_v$v.log 2>&1; echo "tl_decode v$v rc=$?" | tee -a gpurun_out/dec_rc2.log
done
SEGS_DECODE_VARIANT=2 timeout 200 python tools/timeline_mapping.py 8 fused > gpurun_out/tl_mapping_v2.log 2>&1
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 300 $NCU -k regex:decode_forward_v2 -s 3 -c 1 -o gpurun_out/ncu_fwd_v2 python tools/bench_decode.py 200000 --no-oracle > gpurun_out/ncu_fwd.log 2>&1; echo "ncu fwd rc=$?" | tee -a gpurun_out/dec_rc2.log
timeout 300 $NCU -k regex:decode_wgrad_tc -s 2 -c 1 -o gpurun_out/ncu_wgrad_tc python tools/bench_decode.py 200000 --no-oracle > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?" | tee -a gpurun_out/dec_rc2.log
tail -n 4 gpurun_out/dec_all.log
head -n 8 gpurun_out/tl_decode_C3_v1.log gpurun_out/tl_decode_C3_v2.log
grep -E "decode_|span" gpurun_out/tl_mapping_v2.log
ls -la gpurun_out/*.ncu-rep
