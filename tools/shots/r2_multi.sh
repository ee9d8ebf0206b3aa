#!/bin/bash
# Round 2, N-GPU shot: query-sharded parity (world = N) + strong-scaling bench line.  usage: r2_multi.sh N [extra bench args]
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 \
    tests/multi/query_shard_check.py tiny 8 3 > gpurun_out/qshard_tiny_n$N.log 2>&1; echo "qshard tiny exit $?"
tail -4 gpurun_out/qshard_tiny_n$N.log
if [ -z "$SKIP_R50" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29702 \
    tests/multi/query_shard_check.py r50_704x256 8 2 > gpurun_out/qshard_r50_n$N.log 2>&1; echo "qshard r50 exit $?"
tail -4 gpurun_out/qshard_r50_n$N.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29703 \
    bench.py --gpus $N --steps 50 --warmup 5 "$@" > gpurun_out/bench_q_n$N.json 2> gpurun_out/bench_q_n$N.err; echo "bench queries exit $?"
tail -3 gpurun_out/bench_q_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_q_n$N.json').read().strip().splitlines()[-1])
    print('N=$N', 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', (d.get('e2e') or {}).get('value'), 'dp', (d.get('dp_replicas') or {}).get('ms_per_step'))
    print(json.dumps(d.get('exchange'))); print('allgather', json.dumps(d.get('feature_allgather'))[:300]); print('e2e', json.dumps(d.get('e2e'))[:260])
    print('launches', d['launches_per_step'], 'roof', d['roofline'].get('kernel_ms'), d['roofline'].get('frac'))
except Exception as e:
    print('no line', e)
PY
