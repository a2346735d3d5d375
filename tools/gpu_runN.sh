#!/bin/bash
# N GPUs: (2-GPU NCCL tests when N == 2,) the default one-job bench line (configs[2]) incl. the same-run one-GPU baseline
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-2}; TAG=${2:-s10}; shift; shift
free -g | head -2 > $O/${TAG}_n${N}_mem.txt; nproc >> $O/${TAG}_n${N}_mem.txt
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s -k two_gpus > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
  grep -E "passed|failed|Error|archive size" $O/${TAG}_pytest.log | tail -5
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
HARCGPU_ALLOC_LOG=${ALLOC_LOG:-0} timeout 1500 $TR bench.py --gpus $N --steps 5 --warmup 3 "$@" > $O/${TAG}_n${N}.json 2> $O/${TAG}_n${N}.err; echo "bench rc=$?"
python - <<P
import json
try:
    d=json.loads(open("$O/${TAG}_n${N}.json").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, "e2e", d.get("e2e",{}).get("value"), d["verify"]["ok"])
    print(d["stage1"]); print(d["detail"]["device"]["per_step_ms_rank0"], d["detail"]["device"]["cudaMalloc_calls_in_timed_region_rank0"], d["allocator"])
    o=d.get("one_gpu_same_workload"); print(o)
    if o and "ms_per_step" in o: print("strong-scaling efficiency vs the same-run one-GPU time: %.3f" % (o["ms_per_step"]/(d["n_gpus"]*d["ms_per_step"])))
except Exception as e:
    print("ERR", e); print(open("$O/${TAG}_n${N}.err").read()[-2500:])
P
