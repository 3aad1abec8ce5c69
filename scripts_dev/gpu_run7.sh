#!/bin/bash
mkdir -p gpurun_out
# ten-round number at config 3, and the injected-kernel geometry sweep for the superposition form
python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline --no-side --no-injected --philox-rounds 10 > gpurun_out/b_cfg3_r10.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/b_cfg3_r10.json')); print('cfg3 r10', d['ms_per_step'], d['roofline']['frac'])"
for G in "4,4,5" "5,3,5" "4,3,5" "3,4,5" "3,5,5" "2,4,5" "4,4,4" "5,3,5"; do
  MPPI_INJ_GEOM=$G python bench.py --workload cfg3 --steps 6 --warmup 3 --no-cpu-baseline --no-side > gpurun_out/inj_$G.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/inj_$G.json')); print('inj geom $G', d['roofline_injected']['ms_per_launch'], d['roofline_injected']['frac'])"
done
for G in "16,1,48" "8,2,24" "16,1,32" "12,1,36"; do
  MPPI_INJ_GEOM=$G python bench.py --workload cfg5 --steps 6 --warmup 3 --no-cpu-baseline --no-side > gpurun_out/inj5_$G.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/inj5_$G.json')); print('cfg5 inj geom $G', d['roofline_injected']['ms_per_launch'], d['roofline_injected']['frac'])"
done
