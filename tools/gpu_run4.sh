#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
HARCGPU_LAPS=1 timeout 1500 $TR bench.py --gpus 2 --steps 3 --warmup 3 > $O/s4_c2.json 2> $O/s4_c2.err; echo "c2 rc=$?"
python - <<P
import json
try:
    d=json.loads(open("$O/s4_c2.json").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, d.get("e2e",{}).get("value"))
    print(d["verify"]); print(d["laps_ms_rank0"]); print(d.get("one_gpu_same_workload")); print(d["exchange_bytes_per_step"])
except Exception as e:
    print("ERR", e); print(open("$O/s4_c2.err").read()[-2500:])
P
