#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 30 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
SBEV_DENSE_CLUSTER=0 timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu > gpurun_out/bench_nocluster.json 2> gpurun_out/bench_nocluster.err
SBEV_GEMM_IMPL=0 timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu > gpurun_out/bench_noares.json 2> gpurun_out/bench_noares.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 2 --no-graph --skip-cpu > gpurun_out/bench_ncu.log 2>&1
tail -15 gpurun_out/pytest_all.log; tail -3 gpurun_out/bench.err
for f in bench bench_nocluster bench_noares; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['ms_per_step'])"; done
