#!/bin/bash
# one GPU: full GPU test suite, default bench (pipelined e2e), 3 jobs in flight, ncu launch list + full capture of the walk
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/s8_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s8_pytest.log; tail -4 $O/s8_pytest.log
timeout 900 python bench.py --no-cpu-baseline > $O/s8_bench.json 2> $O/s8_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --no-cpu-baseline --pipeline 3 --ingest-reads 0 > $O/s8_bench_p3.json 2> $O/s8_bench_p3.err; echo "bench p3 rc=$?"
for f in bench bench_p3; do python - <<P
import json
try:
    d=json.loads(open("$O/s8_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, d["allocator"])
    e=d["e2e"]; print("  e2e", round(e["value"],1), round(e["ms_per_step"],1), e["what"], "single", round(e["single_job"]["value"],1), e["single_job"]["host_wall_ms"]); print("  pipe", e.get("pipelined_host_wall_ms_per_job"))
except Exception as e:
    print("$f", "ERR", e); print(open("$O/s8_$f.err").read()[-1500:])
P
done
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --ingest-reads 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s8_launches.csv $B > $O/s8_l.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -c 1 -o $O/s8_walk -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --ingest-reads 0 > $O/s8_w.log 2>&1; echo "walk capture rc=$?"
ls -la $O/s8_walk.ncu-rep
