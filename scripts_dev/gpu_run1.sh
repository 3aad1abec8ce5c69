#!/bin/bash
# round-2 GPU call 1: full GPU suite (new full-size parity tests included), then ncu baselines of the round-1 kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/pytest_r2a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2a.log
for W in "cfg3 --k-override 131072:shard131k" "cfg2:cfg2" "cfg5:cfg5"; do
  ARGS="${W%%:*}"; NAME="${W##*:}"
  B="python bench.py --workload $ARGS --steps 6 --warmup 3 --no-cpu-baseline --no-injected"
  $B > gpurun_out/b_$NAME.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:rollout_philox -s 4 -c 1 -f -o gpurun_out/r2a_$NAME $B > gpurun_out/ncu_$NAME.log 2>&1
  echo "$NAME rc=$?" >> gpurun_out/run1_status.txt
done
tail -3 gpurun_out/pytest_r2a.log
