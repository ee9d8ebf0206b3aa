#!/bin/bash
# Round 2, 1-GPU profiling shot: GPU tests, rank-local kernel sequence of the sharded layer (one GPU emulating one rank of N),
# launch lists (ncu gpu__time_duration) of N = 1 and emulated N = 2 / 8.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
LIGHT="--skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e"
for W in 2 4 8; do
  timeout 300 python bench.py --emulate-world $W $LIGHT > gpurun_out/bench_emu$W.json 2> gpurun_out/bench_emu$W.err; echo "emu $W exit $?"
done
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 600 $NCU -c 300 --log-file gpurun_out/launches_n1.csv python bench.py --steps 2 --warmup 3 --no-graph --no-overlap $LIGHT > gpurun_out/ncu_n1.log 2>&1; echo "ncu n1 exit $?"
timeout 600 $NCU -c 300 --log-file gpurun_out/launches_emu8.csv python bench.py --emulate-world 8 --steps 2 --warmup 3 --no-graph --no-overlap $LIGHT > gpurun_out/ncu_emu8.log 2>&1; echo "ncu emu8 exit $?"
timeout 600 $NCU -c 300 --log-file gpurun_out/launches_emu2.csv python bench.py --emulate-world 2 --steps 2 --warmup 3 --no-graph --no-overlap $LIGHT > gpurun_out/ncu_emu2.log 2>&1; echo "ncu emu2 exit $?"
timeout 300 python bench.py --frames 1 --skip-cpu --skip-backbone --skip-e2e > gpurun_out/bench_n1_T1.json 2> gpurun_out/bench_n1_T1.err; echo "bench T1 exit $?"
python - <<PY
import json
for f in ('bench_emu2','bench_emu4','bench_emu8','bench_n1_T1'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms/step', d['ms_per_step'], 'launches', d.get('launches_per_step'), 'gather', (d.get('roofline') or {}).get('kernel_ms'))
        g=d.get('gpu_baseline')
        if g: print('   ', json.dumps(g.get('op'))[:1200])
    except Exception as e:
        print(f, 'no line', e)
PY
