python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo rc=$?
python - <<PY
import json
lines=[l for l in open("gpurun_out/bench_ref_n2.json").read().strip().splitlines() if l.startswith("{")]
print(len(lines), "json lines")
d=json.loads(lines[-1]); print(d["impl"], d["n_gpus"], d["value"], d["e2e"]["value"], d["mapping"]["value"])
PY
