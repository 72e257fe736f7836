# dev: headline stage table (C2) + core parity tests
Q='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["batch"]["value"], {k:v["ms"] for k,v in d["roofline"]["stages"].items()})'
python bench.py --steps 5 --warmup 3 --no-mapping --no-configs --no-e2e --no-cpu-baseline | python -c "$Q"
python -m pytest tests/test_raster_parity_gpu.py tests/test_golden_gpu.py tests/test_aux_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
