#!/bin/bash
# strong-scaling lines of the default workload (run under gpurun --gpus N with N = $1): peer check, then bench at N
N=$1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/peer_check_worker.py 2>&1 | grep "PEER CHECK\|rror" | head -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/scale_r2b_n$N.json 2> gpurun_out/scale_r2b_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/scale_r2b_n$N.json').read().strip().splitlines()[-1]); print('N=$N', round(d['ms_per_step'],5), d['e2e']['latency_ms'], d.get('parity_check'))"
if [ "$N" = "8" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --workload cfg5 --steps 100 --warmup 10 > gpurun_out/scale_r2b_cfg5_n8.json 2> gpurun_out/scale_r2b_cfg5_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/scale_r2b_cfg5_n8.json').read().strip().splitlines()[-1]); print('cfg5 N=8', round(d['ms_per_step'],5), d['value'])"
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts_dev/trace_peers.py 2>&1 | grep "rank" | tail -$((2*N))
