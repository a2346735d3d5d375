#!/bin/bash
# 8 GPUs: BASELINE.json configs[4], 1 B x 100 bp reads from a 3 Gbp genome, ONE job
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-8}
free -g | head -2; nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
HARCGPU_JOB_TIMEOUT_S=120 timeout 1700 $TR bench.py --gpus $N --config 4 --steps 3 --warmup 2 --no-e2e --t1 0 --read-sets-steps 0 > $O/s17_c4_n$N.json 2> $O/s17_c4_n$N.err; echo "rc=$?"
python - <<P
import json
try:
    d=json.loads(open("$O/s17_c4_n$N.json").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, d["verify"])
    print(d["stage1"], d["allocator"], d["detail"]["device"]["per_step_ms_rank0"], d["roofline"]["frac"])
except Exception as e:
    print("ERR", e); print(open("$O/s17_c4_n$N.err").read()[-3000:])
P
nvidia-smi --query-gpu=memory.used --format=csv | head -3
