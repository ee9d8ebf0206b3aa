#!/bin/bash
# 4-GPU box: smoke, scene-parallel and frame-sharded bench at N=4, frame-shard parity at world 4
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port"
$TR 29711 tests/multi/frame_shard_check.py tiny 8 > gpurun_out/frame_shard_w4.log 2>&1; echo "shard check exit $?"
$TR 29712 bench.py --gpus 4 --steps 50 --warmup 5 --skip-cpu > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "n4 exit $?"
$TR 29713 bench.py --gpus 4 --steps 50 --warmup 5 --shard frames --exchange p2p > gpurun_out/bench_frames_p2p_n4.json 2> gpurun_out/bench_frames_p2p_n4.err; echo "frames n4 exit $?"
$TR 29714 bench.py --gpus 4 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n4.json 2> gpurun_out/bench_ref_n4.err; echo "ref n4 exit $?"
tail -3 gpurun_out/smoke.log; grep FRAME_SHARD gpurun_out/frame_shard_w4.log | head -10
for f in bench_n4 bench_frames_p2p_n4; do python -c "import json;d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]);print('$f',d['value'],d['ms_per_step'],d['scaling'])"; done
cut -c1-200 gpurun_out/bench_ref_n4.json
