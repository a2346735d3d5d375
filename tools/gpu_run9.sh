#!/bin/bash
# one GPU: Bloom filter in L2 in front of the stage I tables (bits per key, persisting window)
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --ingest-reads 0"
$B > $O/s9_base.json 2> $O/s9_base.err
HARCGPU_BLOOM1_BITS=5 $B > $O/s9_b5.json 2> $O/s9_b5.err
HARCGPU_BLOOM1_BITS=5 HARCGPU_L2PERSIST=64 $B > $O/s9_b5p.json 2> $O/s9_b5p.err
HARCGPU_BLOOM1_BITS=8 HARCGPU_L2PERSIST=96 $B > $O/s9_b8p.json 2> $O/s9_b8p.err
HARCGPU_BLOOM1_BITS=3 HARCGPU_L2PERSIST=64 $B > $O/s9_b3p.json 2> $O/s9_b3p.err
for f in base b5 b5p b8p b3p; do python - <<P
import json
try:
    d=json.loads(open("$O/s9_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), {k:round(v,2) for k,v in d["phases_ms"].items()}, d["stage1"]["chain_heads"], d["stage1"]["singletons"], round(d["stage1"]["probes_per_read"],2))
except Exception as e: print("$f", "ERR", e); print(open("$O/s9_$f.err").read()[-800:])
P
done
HARCGPU_BLOOM1_BITS=5 HARCGPU_L2PERSIST=64 timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -c 1 -o $O/s9_walk_b5p -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --ingest-reads 0 > $O/s9_w.log 2>&1; echo "walk capture rc=$?"
