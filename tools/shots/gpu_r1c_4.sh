#!/bin/bash
# Session-3 shot 4: evidence for profiles/ -- default bench line (cpu_baseline + backbone record), smoke, ncu launch list, reference arm
mkdir -p gpurun_out
timeout 110 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench exit $?"; tail -2 gpurun_out/bench_c4.err | cut -c1-300
python -c "import json;d=json.load(open('gpurun_out/bench_c4.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_resident_features']['value'], d['roofline']['frac'], d['cpu_baseline'], d['backbone'])" 2>&1 | tail -1 | cut -c1-900
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_c4.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_c4.log
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv \
    python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu --skip-backbone > gpurun_out/bench_ncu_c4.log 2>&1; echo "ncu exit $?"
timeout 60 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c4.json 2> gpurun_out/bench_ref_c4.err; echo "ref arm exit $?"; cut -c1-400 gpurun_out/bench_ref_c4.json
