set -e
Q='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["batch"]["value"], {k:v["ms"] for k,v in d["roofline"]["stages"].items()})'
for v in 10 8 12 6; do
  touch segs_slam_b200/csrc/preprocess.cu; make -C segs_slam_b200/csrc EXTRA=-DPRE_MIN_CTAS=$v -j8 > /dev/null 2>&1
  echo "PRE_MIN_CTAS=$v $(grep -A3 ModeE0 build/segs_raster/preprocess.ptxas.log | grep -o 'Used [0-9]* registers' | head -1)"
  python bench.py --steps 5 --warmup 3 --no-mapping --no-configs --no-e2e --no-cpu-baseline | python -c "$Q"
done
touch segs_slam_b200/csrc/preprocess.cu; make -C segs_slam_b200/csrc -j8 > /dev/null 2>&1
python -m pytest tests/test_raster_parity_gpu.py tests/test_golden_gpu.py tests/test_aux_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
