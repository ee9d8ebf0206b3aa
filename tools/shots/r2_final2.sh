#!/bin/bash
# Round 2 last single-GPU shot on the final tree: the driver's default bench command, the reference arm, and the launch list of the bench command.
mkdir -p gpurun_out
timeout 900 python bench.py --timeline gpurun_out/timeline_n1_final.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
LIGHT="--skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 $LIGHT > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
python - <<PY
import json
for f in ('bench_n1','bench_ref'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', (d.get('e2e') or {}).get('value'), 'launches', d.get('launches_per_step'))
        r=d.get('roofline') or {}
        print('   roofline', r.get('kernel_ms'), 'frac', r.get('frac'), 'traffic', r.get('traffic'))
    except Exception as e:
        print(f, 'no line', e)
PY
