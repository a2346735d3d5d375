#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --ingest-reads 0 > $O/q_bench.json 2> $O/q_bench.err; echo "bench rc=$?"
python - <<P
import json
try:
    d=json.loads(open("$O/q_bench.json").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],2), d["roofline"]["traffic"], d["roofline"]["traffic_note"])
    e=d["e2e"]; print("  e2e", round(e["value"],1), round(e["ms_per_step"],1), e["what"], "single", round(e["single_job"]["value"],1)); print("  pipe", e.get("pipelined_host_wall_ms_per_job"))
except Exception as e:
    print("ERR", e); print(open("$O/q_bench.err").read()[-1500:])
P
