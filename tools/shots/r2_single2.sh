#!/bin/bash
# Round 2, 1-GPU shot: GPU tests, full default bench line (all legs incl. config 4 / 5 e2e), emulated shard ranks with e2e, SASA source profile.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
for W in 2 8; do
  timeout 300 python bench.py --emulate-world $W --skip-cpu --skip-backbone --skip-gpu-baseline > gpurun_out/bench_emu$W.json 2> gpurun_out/bench_emu$W.err; echo "emu $W exit $?"; tail -2 gpurun_out/bench_emu$W.err
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sasa_v3 -s 6 -c 1 -f -o gpurun_out/sasa_m113 python tests/perf/kernel_sweep.py "sasa M113" > gpurun_out/ncu_sasa.log 2>&1; echo "ncu exit $?"
python - <<PY
import json
for f in ('bench_n1','bench_emu2','bench_emu8'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', json.dumps(d.get('e2e'))[:300])
        for k in ('config4_e2e','config5_e2e','backbone'):
            if d.get(k): print('   ',k, json.dumps(d[k])[:500])
        r=d.get('roofline') or {}
        print('   roofline', r.get('kernel_ms'), 'frac', r.get('frac'))
        g=d.get('gpu_baseline')
        if g: print('   gpu_baseline', json.dumps(g)[:1200])
    except Exception as e:
        print(f, 'no line', e)
PY
