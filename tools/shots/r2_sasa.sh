#!/bin/bash
# Round 2 development shot: SASA log2-domain / ftz logits + GEMM epilogue bias prefetch: full GPU suite, N=1 step + timeline, emulated rank of 8.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e --timeline gpurun_out/timeline_n1_s.json > gpurun_out/bench_n1_s.json 2> gpurun_out/bench_n1_s.err; echo "n1 exit $?"
timeout 300 python bench.py --emulate-world 8 --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e --timeline gpurun_out/timeline_emu8_s.json > gpurun_out/bench_emu8_s.json 2> gpurun_out/bench_emu8_s.err; echo "emu8 exit $?"
python - <<'PY'
import json
for f,t in (('bench_n1_s','timeline_n1_s'),('bench_emu8_s','timeline_emu8_s')):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f,'ms/step',d['ms_per_step'],'gemm alone',d['roofline_tensor']['kernel_ms'])
    except Exception as e: print(f,'ERR',e)
    try:
        d=json.load(open('gpurun_out/%s.json'%t))
        print('==',t,'step_us',d.get('step_us'), d.get('error'))
        for k in d.get('kernels',[]): print('%8.2f %7.2f -> %7.2f s%s  %s'%(k['start_us'],k['dur_us'],k['start_us']+k['dur_us'],k['stream'],k['name'][:60]))
    except Exception as e: print(t,'ERR',e)
PY
