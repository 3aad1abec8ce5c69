#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -x --durations=5 > gpurun_out/pytest_r2d.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2d.log
tail -3 gpurun_out/pytest_r2d.log
for W in "131072 100 3" "1024 20 1" "65536 50 2" "1024 30 2 4096" "1048576 100 3"; do
  echo "== trace $W"; python scripts_dev/trace_phases.py $W 2>&1 | tail -9
done > gpurun_out/trace_r2d.txt
for W in "cfg3:cfg3_r7:7" "cfg1:cfg1_r7:7" "cfg2:cfg2_r7:7" "cfg5:cfg5_r7:7" "cfg3 --k-override 131072:shard131k_r7:7"; do
  IFS=: read ARGS NAME R <<< "$W"
  python bench.py --workload $ARGS --steps 20 --warmup 5 --no-cpu-baseline --no-side --philox-rounds $R > gpurun_out/b_$NAME.json 2> gpurun_out/b_$NAME.err
  echo "$NAME rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/b_$NAME.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['latency_ms']['p50'], d['run']['nonzero_weight_frac'], d.get('roofline_injected',{}).get('ms_per_launch'))" 2>&1 | tail -1)"
done
python bench.py --steps 20 --warmup 5 > gpurun_out/b_full_default.json 2> gpurun_out/b_full_default.err; echo "full rc=$?"
