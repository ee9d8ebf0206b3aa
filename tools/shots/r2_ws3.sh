#!/bin/bash
# Round 2 development shot: kernel timelines (CUPTI via torch.profiler) of the N=1 step and of an emulated rank of the 8-GPU step;
# weights-stationary chain re-check after the cheap fixes.
mkdir -p gpurun_out
timeout 300 python bench.py --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e --timeline gpurun_out/timeline_n1.json > gpurun_out/bench_tl_n1.json 2> gpurun_out/bench_tl_n1.err; echo "n1 exit $?"
timeout 300 python bench.py --emulate-world 8 --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e --timeline gpurun_out/timeline_emu8.json > gpurun_out/bench_tl_emu8.json 2> gpurun_out/bench_tl_emu8.err; echo "emu8 exit $?"
timeout 300 python bench.py --no-overlap --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e --timeline gpurun_out/timeline_n1_noov.json > gpurun_out/bench_tl_n1_noov.json 2> gpurun_out/bench_tl_n1_noov.err; echo "n1 no-overlap exit $?"
python - <<'PY'
import json
for f in ('timeline_n1','timeline_emu8','timeline_n1_noov'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print('==',f,'step_us',d.get('step_us'), d.get('error'))
        for k in d.get('kernels',[]): print('%8.2f %7.2f  s%s  %s'%(k['start_us'],k['dur_us'],k['stream'],k['name']))
    except Exception as e: print(f,'ERR',e)
for f in ('bench_tl_n1','bench_tl_emu8','bench_tl_n1_noov'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f,'ms/step',d['ms_per_step'])
    except Exception as e: print(f,'ERR',e)
PY
timeout 600 python -m pytest tests/test_gpu_dense_ws.py -m gpu -q --timeout 120 -x > gpurun_out/ws_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/ws_pytest.log
timeout 400 python tests/perf/ws_sweep.py "M113" "M225" > gpurun_out/ws_sweep.log 2>&1; echo "sweep exit $?"; grep -v stream gpurun_out/ws_sweep.log | tail -30
