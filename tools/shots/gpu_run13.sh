#!/bin/bash
# 2-GPU box, short: frame-sharded parity test + NCCL-exchange bench exit path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
TR="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$TR 29612 bench.py --gpus 2 --steps 50 --warmup 5 --shard frames --exchange nccl > gpurun_out/bench_frames_nccl.json 2> gpurun_out/bench_frames_nccl.err
echo "nccl bench exit $?"
$TR 29613 bench.py --gpus 2 --steps 50 --warmup 5 --shard frames --exchange p2p > gpurun_out/bench_frames_p2p.json 2> gpurun_out/bench_frames_p2p.err
echo "p2p bench exit $?"
grep -E "passed|failed|FRAME_SHARD|Error" gpurun_out/pytest_multi.log | tail -12
for f in bench_frames_p2p bench_frames_nccl; do python -c "import json;d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]);print('$f',d['value'],d['ms_per_step'],d['launches_per_step'])"; done
