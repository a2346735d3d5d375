#!/bin/bash
# 8 GPUs: broadcast of the packed slice by the copy engines against the kernel route (configs[2] and configs[4])
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 8 --steps 5 --warmup 3 > $O/s20_c2_dma.json 2> $O/s20_c2_dma.err; echo "c2 dma rc=$?"
HARCGPU_JOB_BCAST=kernel timeout 900 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --t1 0 > $O/s20_c2_kernel.json 2> $O/s20_c2_kernel.err; echo "c2 kernel rc=$?"
HARCGPU_JOB_TIMEOUT_S=120 timeout 1200 $TR bench.py --gpus 8 --config 4 --steps 3 --warmup 2 --no-e2e --t1 0 > $O/s20_c4_dma.json 2> $O/s20_c4_dma.err; echo "c4 dma rc=$?"
for f in c2_dma c2_kernel c4_dma; do python - <<P
import json
try:
    d=json.loads(open("$O/s20_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, "e2e", d.get("e2e",{}).get("value"), d["verify"]["ok"], d["detail"]["device"]["per_step_ms_rank0"], d["detail"]["device"]["cudaMalloc_calls_in_timed_region_rank0"], d["allocator"])
    o=d.get("one_gpu_same_workload")
    if o and "ms_per_step" in o: print("   one GPU", o["ms_per_step"], "efficiency %.3f" % (o["ms_per_step"]/(d["n_gpus"]*d["ms_per_step"])))
except Exception as e:
    print("$f", "ERR", e); print(open("$O/s20_$f.err").read()[-1500:])
P
done
