#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > gpurun_out/pytest_r2c.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2c.log
tail -3 gpurun_out/pytest_r2c.log
bash scripts_dev/gpu_prof.sh r2_cfg3_philox7 rollout_philox 4 -- --workload cfg3
bash scripts_dev/gpu_prof.sh r2_cfg5_philox7 rollout_philox 4 -- --workload cfg5
bash scripts_dev/gpu_prof.sh r2_cfg2_philox7 rollout_philox 4 -- --workload cfg2
bash scripts_dev/gpu_prof.sh r2_shard131k_philox7 rollout_philox 4 -- --workload cfg3 --k-override 131072
bash scripts_dev/gpu_prof.sh r2_cfg1_philox7 rollout_philox 4 -- --workload cfg1
# injected kernel (superposition form) at config 3: timing only
python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_cfg3_inj.json 2> gpurun_out/b_cfg3_inj.err
python -c "import json; d=json.load(open('gpurun_out/b_cfg3_inj.json')); print('cfg3', d['ms_per_step'], d['roofline']['frac'], d['roofline_injected'])"
python bench.py --workload cfg5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_cfg5_inj.json 2> gpurun_out/b_cfg5_inj.err
python -c "import json; d=json.load(open('gpurun_out/b_cfg5_inj.json')); print('cfg5', d['ms_per_step'], d['roofline']['frac'], d['roofline_injected'])"
du -sh gpurun_out
