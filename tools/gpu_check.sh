#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_stage2_gpu.py tests/test_golden_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > $O/chk_pytest.log 2>&1; echo "pytest rc=$?" >> $O/chk_pytest.log; tail -3 $O/chk_pytest.log
HARCGPU_LAPS=1 timeout 300 python tools/run_shape.py 3e6 100 repeats:20 1 1 > $O/chk_rep.txt 2> $O/chk_rep.err; echo "rep rc=$?"; cat $O/chk_rep.txt
