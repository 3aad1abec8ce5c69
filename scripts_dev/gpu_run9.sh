#!/bin/bash
# round 2, session 2: learner GEMMs on the tensor cores (test), scheduler-balanced warp numbering of the fast kernel (A/B)
mkdir -p gpurun_out
python -m pytest tests/test_train_gpu.py tests/test_philox_variants_gpu.py tests/test_parity_gpu.py tests/test_baseline_configs_gpu.py -x -q -m gpu 2>&1 | tail -5
for v in _build_alt _build; do
  export MPPI_B200_LIB=$PWD/mppi_tf_b200/$v/libmppi_b200.so
  for W in "--k-override 131072:shard" ":full"; do
    IFS=: read ARGS NAME <<< "$W"
    timeout 300 python bench.py $ARGS --steps 100 --warmup 10 --no-cpu-baseline --no-side --no-injected > gpurun_out/ab9_${v}_$NAME.json 2> gpurun_out/ab9_${v}_$NAME.err
    python -c "
import json; d=json.load(open('gpurun_out/ab9_${v}_$NAME.json')); print('$v $NAME', d['ms_per_step'], d['e2e']['latency_ms'])"
  done
  echo "== trace $v shard"; python scripts_dev/trace_phases.py 131072 100 3 2>&1 | tail -10
done
