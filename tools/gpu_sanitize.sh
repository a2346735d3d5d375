#!/bin/bash
# compute-sanitizer over the small GPU cases (golden fixtures incl. the repeat-rich one, stage I/II parity on the smallest
# shapes, one job on several ranks).  Logs go to gpurun_out/; the summaries are copied to profiles/ by hand.
cd "$(dirname "$0")/.."
O=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
T="tests/test_golden_gpu.py tests/test_multigpu_gpu.py::test_one_job_local_ranks"
for tool in memcheck racecheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 --log-file $O/sanitizer_$tool.log --target-processes all \
      python -m pytest $T -m gpu -x -q -k "not w8 and not w4" > $O/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?" >> $O/sanitizer_${tool}_pytest.log
  tail -3 $O/sanitizer_${tool}_pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $O/sanitizer_$tool.log | tail -5
done
