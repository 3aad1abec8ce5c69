#!/bin/bash
# round-2 GPU call: the full GPU suite, then short bench lines of every point-mass config (7 and 10 Philox rounds)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > gpurun_out/pytest_r2b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2b.log
tail -3 gpurun_out/pytest_r2b.log
for W in "cfg3:cfg3_r7:7" "cfg3:cfg3_r10:10" "cfg1:cfg1_r7:7" "cfg2:cfg2_r7:7" "cfg5:cfg5_r7:7" "cfg5:cfg5_r10:10" "cfg3 --k-override 131072:shard131k_r7:7" "cfg3 --k-override 131072:shard131k_r10:10"; do
  IFS=: read ARGS NAME R <<< "$W"
  python bench.py --workload $ARGS --steps 20 --warmup 5 --no-cpu-baseline --philox-rounds $R > gpurun_out/b_$NAME.json 2> gpurun_out/b_$NAME.err
  echo "$NAME rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/b_$NAME.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['latency_ms']['p50'], d['config']['nonzero_weight_frac'], d.get('roofline_injected',{}).get('ms_per_launch'))" 2>&1 | tail -1)"
done
