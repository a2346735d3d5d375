#!/bin/bash
# 8 GPUs: one job (configs[2]) -- default, two walkers per warp, without the Bloom filter
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --t1 0 > $O/s5_n${N}_default.json 2> $O/s5_n${N}_default.err; echo "default rc=$?"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --t1 0 --no-e2e --lanes 16 > $O/s5_n${N}_l16.json 2> $O/s5_n${N}_l16.err; echo "l16 rc=$?"
HARCGPU_JOB_BLOOM=0 timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --t1 0 --no-e2e > $O/s5_n${N}_nobloom.json 2> $O/s5_n${N}_nobloom.err; echo "nobloom rc=$?"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --t1 0 --no-e2e --shard-dicts 0 > $O/s5_n${N}_repl.json 2> $O/s5_n${N}_repl.err; echo "repl rc=$?"
for f in default l16 nobloom repl; do python - <<P
import json
try:
    d=json.loads(open("$O/s5_n${N}_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, d.get("e2e",{}).get("value"), d["verify"]["ok"], d["stage1"]["chain_heads"], d["stage1"]["claim_fails"], d["detail"]["device"]["per_step_ms_rank0"], d["detail"]["device"]["cudaMalloc_calls_in_timed_region_rank0"], d["allocator"])
except Exception as e:
    print("$f", "ERR", e); print(open("$O/s5_n${N}_$f.err").read()[-1500:])
P
done
