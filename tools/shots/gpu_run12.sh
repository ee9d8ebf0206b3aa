#!/bin/bash
# 2-GPU box: full GPU suite (incl. the frame-sharded multi-GPU parity), frame-sharded bench (p2p / nccl), scene-parallel bench
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$TR 29611 bench.py --gpus 2 --steps 50 --warmup 5 --shard frames --exchange p2p > gpurun_out/bench_frames_p2p.json 2> gpurun_out/bench_frames_p2p.err
$TR 29612 bench.py --gpus 2 --steps 50 --warmup 5 --shard frames --exchange nccl > gpurun_out/bench_frames_nccl.json 2> gpurun_out/bench_frames_nccl.err
$TR 29613 bench.py --gpus 2 --steps 50 --warmup 5 --shard frames --exchange p2p --no-graph > gpurun_out/bench_frames_p2p_eager.json 2> gpurun_out/bench_frames_p2p_eager.err
$TR 29614 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
grep -E "passed|failed|error" gpurun_out/pytest_all.log | tail -5
for f in bench_frames_p2p bench_frames_nccl bench_frames_p2p_eager bench_n2 bench; do tail -2 gpurun_out/$f.err | cut -c1-300; python -c "import json;d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]);print('$f',d['value'],d['ms_per_step'],d['launches_per_step'])"; done
