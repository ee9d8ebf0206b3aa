#!/bin/bash
# Round 2 final single-GPU shot: GPU test suite, smoke, default bench line (all legs), T=1 line, reference arm, timelines,
# launch list of the bench command and ncu --set full of the layer's kernels (for profiles/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --timeline gpurun_out/timeline_n1_final.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --frames 1 --skip-cpu --skip-backbone > gpurun_out/bench_n1_T1.json 2> gpurun_out/bench_n1_T1.err; echo "bench T1 exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
for W in 2 4 8; do
  timeout 300 python bench.py --emulate-world $W --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e --timeline gpurun_out/timeline_emu${W}_final.json > gpurun_out/bench_emu$W.json 2> gpurun_out/bench_emu$W.err; echo "emu $W exit $?"
done
LIGHT="--skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 $LIGHT > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"sasa_v3|mix_tma|gemm_bf16|dense_chain_mma|sample_points|sampling4d_c64" -s 33 -c 11 -f -o gpurun_out/layer_kernels python bench.py --steps 2 --warmup 3 --no-graph $LIGHT > gpurun_out/ncu_layer.log 2>&1; echo "ncu layer exit $?"
python - <<PY
import json
for f in ('bench_n1','bench_n1_T1','bench_ref','bench_emu2','bench_emu4','bench_emu8'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', (d.get('e2e') or {}).get('value'), 'launches', d.get('launches_per_step'))
        r=d.get('roofline') or {}
        print('   roofline', r.get('kernel_ms'), 'frac', r.get('frac'), 'traffic', r.get('traffic'))
    except Exception as e:
        print(f, 'no line', e)
PY
