#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
B="timeout 600 python bench.py --steps 50 --warmup 5"
$B --skip-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
$B --skip-cpu --no-overlap > gpurun_out/bench_noov.json 2> gpurun_out/bench_noov.err
grep -E "passed|failed" gpurun_out/pytest_all.log | tail -3; grep -E "^FAILED|Error" gpurun_out/pytest_all.log | head -10; tail -3 gpurun_out/bench.err
for f in bench bench_noov; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['ms_per_step'],d['launches_per_step'])"; done
