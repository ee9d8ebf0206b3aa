#!/bin/bash
# Final 2-GPU evidence: multi-GPU parity tests, scene-parallel and frame-sharded bench lines, reference arm under torchrun
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$TR 29611 bench.py --gpus 2 --steps 50 --warmup 5 --skip-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 exit $?"
$TR 29612 bench.py --gpus 2 --steps 50 --warmup 5 --shard frames --exchange p2p > gpurun_out/bench_frames_p2p.json 2> gpurun_out/bench_frames_p2p.err; echo "frames p2p exit $?"
$TR 29613 bench.py --gpus 2 --steps 50 --warmup 5 --shard frames --exchange nccl > gpurun_out/bench_frames_nccl.json 2> gpurun_out/bench_frames_nccl.err; echo "frames nccl exit $?"
$TR 29614 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref exit $?"
tail -2 gpurun_out/pytest_multi.log
for f in bench_n2 bench_frames_p2p bench_frames_nccl; do python -c "import json;d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]);print('$f',d['value'],d['ms_per_step'],d['scaling'])"; done
cut -c1-160 gpurun_out/bench_ref_n2.json
