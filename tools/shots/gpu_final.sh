#!/bin/bash
# Full measurement pass of the committed state: parity tests, smoke, both bench arms, ncu launch list + full captures.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 50 --warmup 5 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
kill $SMI
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python tests/perf/op_bench.py > gpurun_out/op_bench.log 2>&1
timeout 600 python tests/perf/ref_layer_gpu.py > gpurun_out/ref_layer_gpu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sbev" -s 26 -c 13 -o gpurun_out/prof_layer \
    python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu > gpurun_out/bench_ncu2.log 2>&1
tail -5 gpurun_out/pytest_all.log; cat gpurun_out/smoke.log | tail -2; tail -3 gpurun_out/bench.err
python -c "import json;d=json.load(open('gpurun_out/bench.json'));print('ours',d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e_resident_features']['value'])"
python -c "import json;d=json.load(open('gpurun_out/bench_reference.json'));print('ref',d['value'],d['ms_per_step'],d['cpu_baseline']['cores'])"
