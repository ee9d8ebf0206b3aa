#!/bin/bash
# Round 2, 1-GPU shot: GPU test suite, default bench line (all legs), NHWC-layout bench line, reference arm.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --layout nhwc --skip-cpu --skip-backbone --skip-gpu-baseline > gpurun_out/bench_n1_nhwc.json 2> gpurun_out/bench_n1_nhwc.err; echo "bench nhwc exit $?"
timeout 300 python bench.py --frames 1 --skip-cpu --skip-backbone > gpurun_out/bench_n1_T1.json 2> gpurun_out/bench_n1_T1.err; echo "bench T1 exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
python - <<PY
import json
for f in ('bench_n1','bench_n1_nhwc','bench_n1_T1','bench_ref'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', (d.get('e2e') or {}).get('value'), 'launches', d.get('launches_per_step'))
        r=d.get('roofline') or {}
        print('   roofline', r.get('kernel_ms'), 'frac', r.get('frac'), 'live', r.get('live_tap_fraction'), 'bytes', r.get('bytes'), 'algo', r.get('algorithmic_bytes_upper_bound'))
        ru=d.get('roofline_uniform') or {}
        print('   uniform ', ru.get('kernel_ms'), 'frac', ru.get('frac'))
        g=d.get('gpu_baseline')
        if g: print('   gpu_baseline', json.dumps(g)[:1500])
    except Exception as e:
        print(f, 'no line', e)
PY
