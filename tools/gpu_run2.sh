#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_sort_gpu.py tests/test_stage1_gpu.py -m gpu -q -s > $O/s2b_multi.log 2>&1; echo "multi rc=$?" >> $O/s2b_multi.log
grep -E "passed|failed|Error|error|archive size" $O/s2b_multi.log | tail -30
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --ingest-reads 0"
$B > $O/s2c_base.json 2> $O/s2c_base.err
python - <<P
import json
d=json.loads(open("$O/s2c_base.json").read().strip().splitlines()[-1])
print(round(d["value"],1), d["phases_ms"], d["stage1"])
P
