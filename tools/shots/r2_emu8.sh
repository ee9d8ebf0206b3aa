#!/bin/bash
# Round 2 development shot: in-situ A/B of the sharded layer's knobs on ONE emulated rank of the 8-GPU (and 4-GPU) run.
mkdir -p gpurun_out
run() { # world, name, env..., -- extra args
  local W=$1 name=$2; shift 2
  timeout 200 env "$@" python bench.py --emulate-world $W --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e > gpurun_out/bench_e${W}_$name.json 2> gpurun_out/bench_e${W}_$name.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_e${W}_$name.json').read().strip().splitlines()[-1]);print('emu$W $name ms/step %.4f' % d['ms_per_step'])" 2>/dev/null || echo "emu$W $name FAILED"
}
run 8 base X=1
for K in 36 48 64 72 96; do run 8 splitk$K SBEV_QSHARD_SPLIT_K=$K; done
run 8 kq4 SBEV_SASA_KQ=4
run 8 kq8 SBEV_SASA_KQ=8
run 8 gather4 SBEV_GATHER_VARIANT=4
run 8 gather2 SBEV_GATHER_VARIANT=2
run 8 gemm3 SBEV_GEMM_IMPL=3
run 8 nopdl SBEV_PDL=0
run 4 base X=1
for K in 36 48 72 96 128; do run 4 splitk$K SBEV_QSHARD_SPLIT_K=$K; done
run 2 base X=1
for K in 18 24 36 72; do run 2 splitk$K SBEV_QSHARD_SPLIT_K=$K; done
