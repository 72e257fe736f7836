#!/bin/bash
# Dev tool (GPU box): rasterizer parity (bit-exact binning) and the per-stage times of the M1 path.
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 200 --timeout-method thread"
timeout 600 $PT tests/test_raster_parity_gpu.py tests/test_golden_gpu.py tests/test_aux_parity_gpu.py tests/test_capi_cpp_gpu.py > gpurun_out/raster_tests.log 2>&1; echo "raster tests rc=$?"; tail -n 1 gpurun_out/raster_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-mapping --no-configs --no-e2e --no-cpu-baseline 2> gpurun_out/bench_quick.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('M1', d['value'], 'it/s', d['ms_per_view'], 'ms/view; batch', d['batch']['value'])
print({k:v['ms'] for k,v in d['roofline']['stages'].items()})"
