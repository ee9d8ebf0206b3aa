#!/bin/bash
# Round 2 closing shot: layer / drop-in tests + the driver's default bench command on the final tree.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layer.py tests/test_gpu_dropin.py -m gpu -q --timeout 300 > gpurun_out/pytest_layer.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_layer.log
timeout 900 python bench.py --timeline gpurun_out/timeline_n1_final.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --frames 1 --skip-cpu --skip-backbone > gpurun_out/bench_n1_T1.json 2> gpurun_out/bench_n1_T1.err; echo "bench T1 exit $?"
python - <<PY
import json
for f in ('bench_n1','bench_n1_T1'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', (d.get('e2e') or {}).get('value'), 'launches', d.get('launches_per_step'), 'resident', (d.get('e2e_resident_features') or {}).get('ms_per_step'))
    except Exception as e:
        print(f, 'no line', e)
PY
