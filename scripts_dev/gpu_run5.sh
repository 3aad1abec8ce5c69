#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/pytest_r2e.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2e.log
tail -3 gpurun_out/pytest_r2e.log
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_r2e.log | head -20
python bench.py --steps 20 --warmup 5 > gpurun_out/b_full_default.json 2> gpurun_out/b_full_default.err; echo "full rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/b_full_default.json'))
print(d['ms_per_step'], d['roofline']['frac'], d['roofline_injected']['frac'], d['dense_weights']['ms_per_step'])
for k,v in d['other_configs'].items(): print(k, round(v['ms_per_step'],5), round(v['roofline']['frac'],3), v['nonzero_weight_frac'], v.get('dense_weights'))
"
