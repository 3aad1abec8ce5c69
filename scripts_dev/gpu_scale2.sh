#!/bin/bash
# multi-GPU bench lines (run under gpurun --gpus 8): N = 8, 4, 2 on config 3 + config 5 at 8, each with parity_check
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/scale_gpus.txt
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/scale_r2_n$N.json 2> gpurun_out/scale_r2_n$N.err
  echo "N=$N rc=$? $(python -c "import json; d=json.load(open('gpurun_out/scale_r2_n$N.json')); print(d['ms_per_step'], d['e2e']['latency_ms']['p50'], d['parity_check'])" 2>&1 | tail -1)"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --workload cfg5 --steps 50 --warmup 10 > gpurun_out/scale_r2_cfg5_n8.json 2> gpurun_out/scale_r2_cfg5_n8.err
echo "cfg5 N=8 rc=$? $(python -c "import json; d=json.load(open('gpurun_out/scale_r2_cfg5_n8.json')); print(d['ms_per_step'], d['parity_check'])" 2>&1 | tail -1)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/peer_check_worker.py > gpurun_out/peer_check_n2.log 2>&1; echo "peer check rc=$?"; tail -2 gpurun_out/peer_check_n2.log
python -m pytest tests/test_peer_exchange_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
