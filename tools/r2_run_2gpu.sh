#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_final_2gpu.json 2> gpurun_out/r2_final_2gpu.err
tail -c 600 gpurun_out/r2_final_2gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_final_2gpu.json").read().strip().splitlines()[-1])
print("2gpu", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["n_gpus"], d["config"]["global_batch"])
PY
