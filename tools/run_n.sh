#!/bin/bash
# usage: bash tools/run_n.sh N  -> gpurun_out/bench_n${N}_l4.json (bench.py under torchrun, one rank per GPU)
N=$1
if [ "$N" = "1" ]; then
  python bench.py > gpurun_out/bench_n1_l4.json 2> gpurun_out/bench_n1_l4.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_l4.json 2> gpurun_out/bench_n${N}_l4.err
fi
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n${N}_l4.json").read().strip().splitlines()[-1])
print("N=${N}", d["value"], d["e2e"]["value"], d["mapping"]["value"], d["mapping"]["losses"], d["clocks"]["sm_mhz"])
PY
