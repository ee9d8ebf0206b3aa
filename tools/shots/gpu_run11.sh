#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
B="timeout 600 python bench.py --steps 50 --warmup 5"
$B --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
$B --skip-cpu --split-k 16 > gpurun_out/bench_sk16.json 2> gpurun_out/bench_sk16.err
$B --skip-cpu --no-tma-params > gpurun_out/bench_notma.json 2> gpurun_out/bench_notma.err
$B --skip-cpu --opt dense_cluster=2 > gpurun_out/bench_dc2.json 2> gpurun_out/bench_dc2.err
$B --skip-cpu --opt dense_cluster=4 > gpurun_out/bench_dc4.json 2> gpurun_out/bench_dc4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dc4.csv \
    python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu --opt dense_cluster=4 > gpurun_out/bench_ncu_dc4.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_all.log | tail -3; tail -3 gpurun_out/bench.err
for f in bench bench_sk16 bench_notma bench_dc2 bench_dc4; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['ms_per_step'],d['launches_per_step'])"; done
