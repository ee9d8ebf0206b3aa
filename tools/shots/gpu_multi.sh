#!/bin/bash
# N-GPU check of the driver's launch recipe (one rank per GPU over NCCL), both arms.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
tail -2 gpurun_out/bench_n$N.json | cut -c1-600; tail -3 gpurun_out/bench_n$N.err; tail -1 gpurun_out/bench_ref_n$N.json | cut -c1-300
