for v in _build_old _build; do
  export MPPI_B200_LIB=$PWD/mppi_tf_b200/$v/libmppi_b200.so
  for W in "--workload cfg4:cfg4" "--workload auv:auv"; do
    IFS=: read ARGS NAME <<< "$W"
    timeout 300 python bench.py $ARGS --steps 50 --warmup 10 --no-cpu-baseline --no-side --no-injected > gpurun_out/ab4_${v}_$NAME.json 2> gpurun_out/ab4_${v}_$NAME.err
    python -c "
import json; d=json.load(open('gpurun_out/ab4_${v}_$NAME.json')); print('$v $NAME', round(d['ms_per_step'],5), d['e2e']['latency_ms']['p50'], d['clocks'])"
  done
done
