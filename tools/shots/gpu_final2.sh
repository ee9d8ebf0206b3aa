#!/bin/bash
# Final single-GPU evidence run of round 1: full GPU suite, bench (both arms), clocks, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
( while true; do nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw --format=csv,noheader >> gpurun_out/clocks.csv; sleep 0.2; done ) &
CLK=$!
timeout 600 python bench.py --steps 50 --warmup 5 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
kill $CLK
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dense_chain|mix_|sasa_v3|persistent|sampling4d|reduce_ln|sample_points" -s 22 -c 11 -o gpurun_out/prof_layer \
    python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu > gpurun_out/bench_ncu2.log 2>&1
timeout 300 python tests/perf/op_bench.py > gpurun_out/op_bench.json 2> gpurun_out/op_bench.err
grep -E "passed|failed" gpurun_out/pytest_all.log | tail -3; tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/bench.err
python -c "import json;d=json.load(open('gpurun_out/bench.json'));print('bench',d['value'],d['ms_per_step'],d['launches_per_step'],d['e2e']['value'],d['e2e_resident_features']['value'],d['roofline']['frac'],d['roofline_tensor']['frac'],d['cpu_baseline']['value'])"
cut -c1-300 gpurun_out/bench_ref.json
