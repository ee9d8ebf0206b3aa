#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120 -x -k "bf16x3_is_fp32_grade or split_output" > gpurun_out/pytest_pair.log 2>&1
echo "pair pytest exit $?" >> gpurun_out/pytest_pair.log
tail -3 gpurun_out/pytest_pair.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
B="timeout 600 python bench.py --steps 50 --warmup 5"
$B --skip-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
$B --skip-cpu --opt gemm_impl=3 > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err
$B --skip-cpu --no-overlap > gpurun_out/bench_noov.json 2> gpurun_out/bench_noov.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu > gpurun_out/bench_ncu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_all.log | tail -3; tail -3 gpurun_out/bench.err
for f in bench bench_single bench_noov; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['ms_per_step'],d['launches_per_step'],d['roofline_tensor']['achieved'],d['roofline_tensor']['kernel_ms'])"; done
