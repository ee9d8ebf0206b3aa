#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
B="timeout 600 python bench.py --steps 50 --warmup 5"
$B --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
$B --skip-cpu --opt pdl=0 > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err
$B --skip-cpu --no-graph > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err
grep -E "passed|failed" gpurun_out/pytest_all.log | tail -3; tail -3 gpurun_out/bench.err
for f in bench bench_nopdl bench_eager; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['ms_per_step'],d['launches_per_step'],d['e2e_resident_features']['value'])"; done
