#!/bin/bash
# Round 2 development shot: in-situ A/B of kernel-variant options on the N=1 step (the step time is what counts: isolated kernel
# times mislead where two kernels share the L2 throughput cap).
mkdir -p gpurun_out
run() { # name, extra args
  local name=$1; shift
  timeout 200 python bench.py --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e "$@" > gpurun_out/bench_opt_$name.json 2> gpurun_out/bench_opt_$name.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_opt_$name.json').read().strip().splitlines()[-1]);print('$name ms/step %.4f' % d['ms_per_step'])" 2>/dev/null || echo "$name FAILED"
}
run base
run sasa_kq8 --opt sasa_kq=8
run gather2 --opt gather_variant=2
run gather4 --opt gather_variant=4
run gather5 --opt gather_variant=5
run gather3 --opt gather_variant=3
run mix_order1 --opt mix_order=1
run gemm4 --opt gemm_impl=4
run gemm3 --opt gemm_impl=3
run splitk36 --split-k 36
run splitk12 --split-k 12
run nopdl --opt pdl=0
run base2
