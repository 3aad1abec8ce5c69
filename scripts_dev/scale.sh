# strong-scaling run of the default workload on one box: N = 1, 2, 4, 8 (run under gpurun --gpus 8)
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-injected > gpurun_out/scale_r1_n1.json 2> gpurun_out/scale_n1.err
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/scale_r1_n$n.json 2> gpurun_out/scale_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 100 --warmup 10 --exchange nccl > gpurun_out/scale_r1_n8_nccl.json 2> gpurun_out/scale_n8_nccl.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --workload cfg5 --steps 100 --warmup 10 > gpurun_out/scale_r1_cfg5_n8.json 2> gpurun_out/scale_cfg5_n8.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/peer_check_worker.py 2>&1 | grep "PEER CHECK"
for f in gpurun_out/scale_r1_*.json; do python -c "
import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], d['config']['exchange'], round(d['ms_per_step'],4), d['value'], d['e2e']['latency_ms']['p50'])"; done
