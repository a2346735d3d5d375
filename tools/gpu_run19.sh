#!/bin/bash
# 2 GPUs: per-GPU shard size of configs[4] at N = 8 (125 M reads per GPU) with the lap timers of the dictionary build
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
HARCGPU_LAPS=1 timeout 1200 $TR bench.py --gpus 2 --config 4 --reads 250e6 --genome 750e6 --steps 3 --warmup 2 --no-e2e --t1 0 > $O/s19.json 2> $O/s19.err; echo "rc=$?"
python - <<P
import json
try:
    d=json.loads(open("$O/s19.json").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, d["verify"]["ok"])
    print({k:round(v,2) for k,v in d["laps_ms_rank0"].items()}, d["allocator"])
except Exception as e:
    print("ERR", e); print(open("$O/s19.err").read()[-2000:])
P
