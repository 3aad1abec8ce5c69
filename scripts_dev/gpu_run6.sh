#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_r2f.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2f.log
tail -2 gpurun_out/pytest_r2f.log
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_r2f.log | head
bash scripts_dev/bounds_check.sh
rm -rf mppi_tf_b200/_build_dbg
bash scripts_dev/gpu_prof.sh r2_cfg3_philox7 rollout_philox 4 -- --workload cfg3 --no-side
bash scripts_dev/gpu_prof.sh r2_cfg5_philox7 rollout_philox 4 -- --workload cfg5 --no-side
bash scripts_dev/gpu_prof.sh r2_cfg2_philox7 rollout_philox 4 -- --workload cfg2 --no-side
bash scripts_dev/gpu_prof.sh r2_shard131k_philox7 rollout_philox 4 -- --workload cfg3 --k-override 131072
bash scripts_dev/gpu_prof.sh r2_cfg1_philox7 rollout_philox 4 -- --workload cfg1 --no-side
# the injected-noise kernel at config 3 (bench launches it after the Philox loops)
B="python bench.py --workload cfg3 --steps 6 --warmup 3 --no-cpu-baseline --no-side"
$B > gpurun_out/plain_r2_cfg3_injected.json 2> gpurun_out/plain_r2_cfg3_injected.err && \
ncu --set full --clock-control none --import-source on -k regex:rollout_injected -s 2 -c 1 -f -o gpurun_out/prof_r2_cfg3_injected $B > gpurun_out/ncu_r2_cfg3_injected.log 2>&1
python profiles/summarize_ncu.py gpurun_out/prof_r2_cfg3_injected.ncu-rep > gpurun_out/r2_cfg3_injected_ncu_summary.txt 2>&1
ncu -i gpurun_out/prof_r2_cfg3_injected.ncu-rep --page source --csv > gpurun_out/r2_cfg3_injected_source.csv 2>/dev/null
rm -f gpurun_out/prof_r2_cfg3_injected.ncu-rep
# launch list of the default command
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-side"
$B > gpurun_out/plain_default.json 2> gpurun_out/plain_default.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg3.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
du -sh gpurun_out
