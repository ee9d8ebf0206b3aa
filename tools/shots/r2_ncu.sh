#!/bin/bash
# Round 2, 1-GPU profiling shot for profiles/: launch list of the bench command, ncu --set full of the gather (T = 8, T = 1) and of
# the op-boundary kernel (uniform), plus the GPU tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
LIGHT="--skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 $LIGHT > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sampling4d_c64 -s 8 -c 1 -f -o gpurun_out/gather_T8 python bench.py --steps 2 --warmup 3 $LIGHT > gpurun_out/ncu_g8.log 2>&1; echo "ncu gather T8 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sampling4d_c64 -s 8 -c 1 -f -o gpurun_out/gather_T1 python bench.py --frames 1 --steps 2 --warmup 3 $LIGHT > gpurun_out/ncu_g1.log 2>&1; echo "ncu gather T1 exit $?"
timeout 900 ncu --set full --clock-control none -k regex:msmv_fwd_c64 -s 2 -c 1 -f -o gpurun_out/op_uniform_T8 python bench.py --steps 2 --warmup 3 $LIGHT > gpurun_out/ncu_op.log 2>&1; echo "ncu op exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"sasa_v3|mix_tma|gemm_bf16|dense_chain_mma|peer_exchange|sample_points" -s 30 -c 12 -f -o gpurun_out/layer_kernels python bench.py --steps 2 --warmup 3 --no-graph $LIGHT > gpurun_out/ncu_layer.log 2>&1; echo "ncu layer exit $?"
timeout 300 python tests/perf/kernel_sweep.py "sasa" "gather v6" > gpurun_out/sweep_sasa.log 2>&1; grep "sasa\|gather" gpurun_out/sweep_sasa.log
ls -la gpurun_out/*.ncu-rep
