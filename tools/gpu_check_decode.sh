#!/bin/bash
# Dev tool (GPU box): the decode kernels of both variants — parity tests in separate processes (a hang costs one
# process), then CUPTI timelines at C3 and of a mapping view.  Logs under gpurun_out/.
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 150 --timeout-method thread"
timeout 400 $PT tests/test_decode_gpu.py -k "forward_parity or ragged or empty" > gpurun_out/dec_fwd.log 2>&1; echo "fwd rc=$?" | tee -a gpurun_out/dec_rc.log
SEGS_DECODE_WGRAD=1 timeout 300 $PT tests/test_decode_gpu.py -k "backward_parity or variants_agree" > gpurun_out/dec_bwd_wg1.log 2>&1; echo "bwd(wgrad v1) rc=$?" | tee -a gpurun_out/dec_rc.log
timeout 300 $PT tests/test_decode_gpu.py -k "backward_parity or variants_agree" > gpurun_out/dec_bwd.log 2>&1; echo "bwd rc=$?" | tee -a gpurun_out/dec_rc.log
timeout 400 $PT tests/test_decode_gpu.py -k "C3_size" tests/test_reference_chain_gpu.py > gpurun_out/dec_chain.log 2>&1; echo "C3+chain rc=$?" | tee -a gpurun_out/dec_rc.log
for v in 1 2; do
  SEGS_DECODE_VARIANT=$v timeout 200 python tools/timeline_decode.py > gpurun_out/tl_decode_C3_v$v.log 2>&1; echo "tl_decode v$v rc=$?" | tee -a gpurun_out/dec_rc.log
  SEGS_DECODE_VARIANT=$v timeout 200 python tools/timeline_mapping.py 8 fused > gpurun_out/tl_mapping_v$v.log 2>&1; echo "tl_mapping v$v rc=$?" | tee -a gpurun_out/dec_rc.log
done
tail -3 gpurun_out/dec_fwd.log gpurun_out/dec_bwd_wg1.log gpurun_out/dec_bwd.log gpurun_out/dec_chain.log
head -12 gpurun_out/tl_decode_C3_v1.log gpurun_out/tl_decode_C3_v2.log
