#!/bin/bash
# one GPU: walk kernel bounded for 9 / 10 blocks per SM instead of 8 (-DWALK_MB, library rebuilt on the box only)
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --ingest-reads 0"
for mb in 9 10; do
  HARC_CUFLAGS=-DWALK_MB=$mb python harc_b200/build.py -f > /dev/null 2>&1
  $B > $O/mb$mb.json 2> $O/mb$mb.err
  python - <<P
import json
try:
    d=json.loads(open("$O/mb$mb.json").read().strip().splitlines()[-1])
    print("WALK_MB=$mb", round(d["value"],1), {k:round(v,2) for k,v in d["phases_ms"].items()}, d["stage1"]["chain_heads"], d["stage1"]["singletons"])
except Exception as e: print("ERR", e); print(open("$O/mb$mb.err").read()[-800:])
P
done
