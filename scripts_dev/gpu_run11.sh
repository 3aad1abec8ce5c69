#!/bin/bash
# A/B of the library in _build_alt (previous) against _build (current): event-timed updates and phase traces
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for v in _build_alt _build; do
  export MPPI_B200_LIB=$PWD/mppi_tf_b200/$v/libmppi_b200.so
  for W in "--k-override 131072:shard" ":full" "--workload cfg2:cfg2" "--workload cfg1:cfg1" "--workload cfg5:cfg5"; do
    IFS=: read ARGS NAME <<< "$W"
    timeout 300 python bench.py $ARGS --steps 100 --warmup 10 --no-cpu-baseline --no-side --no-injected > gpurun_out/ab11_${v}_$NAME.json 2> gpurun_out/ab11_${v}_$NAME.err
    python -c "
import json; d=json.load(open('gpurun_out/ab11_${v}_$NAME.json')); print('$v $NAME', round(d['ms_per_step'],5), d['e2e']['latency_ms'])"
  done
  for W in "131072 100 3" "65536 50 2"; do echo "== trace $v $W"; python scripts_dev/trace_phases.py $W 2>&1 | tail -14 | grep "published\|merged\|applied\|wsum"; done
done
