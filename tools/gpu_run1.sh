#!/bin/bash
# round 2, GPU run 1: parity of the single-wave walk + L2 experiments
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/s2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s2a_pytest.log
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --ingest-reads 0"
$B > $O/s2a_base.json 2> $O/s2a_base.err
HARCGPU_L2FETCH=32 $B > $O/s2a_f32.json 2> $O/s2a_f32.err
HARCGPU_L2PERSIST=16 $B > $O/s2a_p16.json 2> $O/s2a_p16.err
HARCGPU_L2FETCH=32 HARCGPU_L2PERSIST=16 $B > $O/s2a_f32p16.json 2> $O/s2a_f32p16.err
tail -2 $O/s2a_pytest.log
for f in base f32 p16 f32p16; do python - <<P
import json
try:
    d=json.loads(open("$O/s2a_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), d["phases_ms"], d["stage1"]["chain_heads"], d["stage1"]["singletons"], d["stage1"]["probes_per_read"], d["stage1"]["claim_fails"])
except Exception as e: print("$f", "ERR", e)
P
done
