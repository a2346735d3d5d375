#!/bin/bash
# 2 GPUs: the one-job path over real NVLink peer memory (CUDA IPC, barrier kernels, NCCL glue)
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi -L > $O/s3_gpus.txt
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s -k two_gpus > $O/s3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s3_pytest.log
grep -E "passed|failed|Error|error|archive size" $O/s3_pytest.log | tail
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --config 1 --steps 3 --warmup 3 --t1 0 > $O/s3_c1_sd1.json 2> $O/s3_c1_sd1.err; echo "c1 sd1 rc=$?"
timeout 900 $TR bench.py --gpus 2 --config 1 --steps 3 --warmup 3 --t1 0 --shard-dicts 0 --no-e2e > $O/s3_c1_sd0.json 2> $O/s3_c1_sd0.err; echo "c1 sd0 rc=$?"
timeout 1200 $TR bench.py --gpus 2 --steps 3 --warmup 2 > $O/s3_c2.json 2> $O/s3_c2.err; echo "c2 rc=$?"
for f in c1_sd1 c1_sd0 c2; do python - <<P
import json
try:
    d=json.loads(open("$O/s3_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, d.get("e2e",{}).get("value"), d["verify"], d["stage1"], d.get("one_gpu_same_workload"))
except Exception as e:
    print("$f", "ERR", e); print(open("$O/s3_$f.err").read()[-1500:])
P
done
