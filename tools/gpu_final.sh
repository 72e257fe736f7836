#!/bin/bash
# Round-end evidence on one B200: GPU tests, smoke, both bench arms, then ncu captures of the decode kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -n 2 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench ours rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 300 $NCU -k regex:decode_ -s 26 -c 4 -o gpurun_out/ncu_decode_C3 python tools/bench_decode.py 200000 --no-oracle > gpurun_out/ncu_decode_C3.log 2>&1; echo "ncu decode C3 rc=$?"
timeout 300 $NCU -k regex:decode_ -s 6 -c 6 -o gpurun_out/ncu_decode_view python tools/one_view_mapping.py 2 > gpurun_out/ncu_decode_view.log 2>&1; echo "ncu decode view rc=$?"
head -c 600 gpurun_out/bench_ours.json; echo
