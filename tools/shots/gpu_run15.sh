#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
B="timeout 600 python bench.py --steps 50 --warmup 5"
$B --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
$B --skip-cpu --no-tma-params > gpurun_out/bench_notma.json 2> gpurun_out/bench_notma.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dense_chain|mix_|sasa_v3|persistent|sampling4d|reduce_ln|sample_points" -s 22 -c 11 -o gpurun_out/prof_layer \
    python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu > gpurun_out/bench_ncu2.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_all.log | tail -3; tail -3 gpurun_out/bench.err
for f in bench bench_notma; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['ms_per_step'],d['launches_per_step'],d['e2e_resident_features']['value'])"; done
