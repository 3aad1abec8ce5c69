#!/bin/bash
mkdir -p gpurun_out
for W in "cfg1:cfg1_r7:7" "cfg2:cfg2_r7:7" "cfg5:cfg5_r7:7"; do
  IFS=: read ARGS NAME R <<< "$W"
  python bench.py --workload $ARGS --steps 30 --warmup 5 --no-cpu-baseline --no-side --no-injected --philox-rounds $R > gpurun_out/b_$NAME.json 2> gpurun_out/b_$NAME.err
  echo "$NAME rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/b_$NAME.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['latency_ms']['p50'])" 2>&1 | tail -1)"
done
for W in "1024 20 1" "65536 50 2"; do echo "== trace $W"; python scripts_dev/trace_phases.py $W 2>&1 | tail -8; done
