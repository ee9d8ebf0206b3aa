#!/bin/bash
# Round 2 development shot: GPU tests, chain kernels with / without the packed weight stream, stage-wise parity attribution.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
SBEV_DENSE_PACK=1 timeout 600 python tests/perf/kernel_sweep.py chainA chainB "ffn M" "cls M" "reg M" > gpurun_out/sweep_pack1.log 2>&1; echo "sweep pack1 exit $?"
cp gpurun_out/kernel_sweep.json gpurun_out/kernel_sweep_pack1.json
SBEV_DENSE_PACK=0 timeout 600 python tests/perf/kernel_sweep.py chainA chainB "ffn M" "cls M" "reg M" > gpurun_out/sweep_pack0.log 2>&1; echo "sweep pack0 exit $?"
cp gpurun_out/kernel_sweep.json gpurun_out/kernel_sweep_pack0.json
paste <(grep "ns0" gpurun_out/sweep_pack1.log) <(grep "ns0" gpurun_out/sweep_pack0.log | awk '{print $NF}')
timeout 900 python tests/perf/parity_stages.py > gpurun_out/parity_stages.log 2>&1; echo "parity exit $?"; cat gpurun_out/parity_stages.log
