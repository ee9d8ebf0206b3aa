#!/bin/bash
# Round 2 development shot: GPU tests; chain / SASA sweeps; ncu source-level capture of one chain launch.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python tests/perf/kernel_sweep.py chainA chainB "ffn M" "cls M" "reg M" "sasa" > gpurun_out/sweep3.log 2>&1; echo "sweep exit $?"
grep "ns0\|sasa" gpurun_out/sweep3.log
timeout 600 ncu --set full --import-source on -k regex:dense_chain_mma -s 6 -c 1 -f -o gpurun_out/chain_cls python tests/perf/kernel_sweep.py "cls M900 ns0" > gpurun_out/ncu_chain.log 2>&1; echo "ncu exit $?"
for W in 2 8; do
  timeout 300 python bench.py --emulate-world $W --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e > gpurun_out/bench_emu$W.json 2> gpurun_out/bench_emu$W.err; echo "emu $W exit $?"
  python -c "import json;d=json.loads(open('gpurun_out/bench_emu$W.json').read().strip().splitlines()[-1]);print('emu$W ms/step',d['ms_per_step'],'launches',d['launches_per_step'])"
done
timeout 300 python bench.py --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e > gpurun_out/bench_n1_quick.json 2> gpurun_out/bench_n1_quick.err; echo "n1 exit $?"
python -c "import json;d=json.loads(open('gpurun_out/bench_n1_quick.json').read().strip().splitlines()[-1]);print('n1 ms/step',d['ms_per_step'],'launches',d['launches_per_step'])"
