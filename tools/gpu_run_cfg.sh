#!/bin/bash
# one GPU: BASELINE.json configs[3] (100M x 250bp, -p) and configs[0] (1M error-free reads) through bench.py --config
cd "$(dirname "$0")/.."
O=gpurun_out
free -g | head -2
timeout 1500 python bench.py --config 3 --no-cpu-baseline --steps 3 --warmup 2 --ingest-reads 0 --pipeline 1 > $O/s13_c3.json 2> $O/s13_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --config 0 --steps 5 --warmup 3 --ingest-reads 0 > $O/s13_c0.json 2> $O/s13_c0.err; echo "c0 rc=$?"
for f in c3 c0; do python - <<P
import json
try:
    d=json.loads(open("$O/s13_$f.json").read().strip().splitlines()[-1])
    print("$f", d["metric"], round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, d["allocator"])
    print("  ", d["stage1"], d["stage2"], d["config"]["clean_reads"], d["config"]["reads_with_N"])
    e=d["e2e"]; print("  e2e", round(e["value"],1), e["what"], "single", round(e["single_job"]["value"],1), e["single_job"]["host_wall_ms"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"])
    print("  roof", d["roofline"]["frac"], d["roofline"]["algorithmic_bytes_per_clean_read"], "bits", d.get("bits_per_base"), "cpu", d.get("cpu_baseline"))
except Exception as e:
    print("$f", "ERR", e); print(open("$O/s13_$f.err").read()[-1500:])
P
done
# racecheck again after the __syncwarp in the sort's ranking loop (one golden fixture is enough to run every sort route)
timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --print-limit 5 --log-file $O/sanitizer_racecheck2.log python -m pytest tests/test_golden_gpu.py -m gpu -x -q -k "L100_RC or repeats" > $O/sanitizer_racecheck2_pytest.log 2>&1
tail -2 $O/sanitizer_racecheck2_pytest.log; grep -E "RACECHECK SUMMARY" $O/sanitizer_racecheck2.log
