#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
HARCGPU_ALLOC_LOG=1 timeout 900 $TR bench.py --gpus $N --steps 6 --warmup 3 --t1 0 --no-e2e > $O/s6_n${N}.json 2> $O/s6_n${N}.err; echo "rc=$?"
grep "harcgpu dev 0" $O/s6_n${N}.err | tail -40
python - <<P
import json
d=json.loads(open("$O/s6_n${N}.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, d["detail"]["device"]["per_step_ms_rank0"], d["detail"]["device"]["cudaMalloc_calls_in_timed_region_rank0"], d["detail"]["device"]["phases_ms_per_step_rank0"]["finalize"])
P
