#!/bin/bash
# round 2, session 2 closing run: full default bench line, reference arm, refreshed ncu captures, launch list
mkdir -p gpurun_out
S=$(date +%s)
python bench.py > gpurun_out/bench_r2b_full.json 2> gpurun_out/bench_r2b_full.err; echo "bench rc=$? $(( $(date +%s) - S )) s"
S=$(date +%s)
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2b_reference.json 2> gpurun_out/bench_r2b_reference.err; echo "reference rc=$? $(( $(date +%s) - S )) s"
bash scripts_dev/gpu_prof.sh r2b_cfg3_philox7 rollout_philox 4 -- --workload cfg3 --no-side
bash scripts_dev/gpu_prof.sh r2b_shard131k_philox7 rollout_philox 4 -- --workload cfg3 --k-override 131072 --no-side
bash scripts_dev/gpu_prof.sh r2b_cfg2_philox7 rollout_philox 4 -- --workload cfg2 --no-side
bash scripts_dev/gpu_prof.sh r2b_cfg5_philox7 rollout_philox 4 -- --workload cfg5 --no-side
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-side"
$B > gpurun_out/plain_default.json 2> gpurun_out/plain_default.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_cfg3.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python scripts_dev/trace_phases.py 131072 100 3 2>&1 | tail -14 > gpurun_out/trace_r2b.txt
for W in "1048576 100 3" "65536 50 2" "1024 20 1" "1024 30 2 4096"; do echo "== trace $W" >> gpurun_out/trace_r2b.txt; python scripts_dev/trace_phases.py $W 2>&1 | tail -14 >> gpurun_out/trace_r2b.txt; done
du -sh gpurun_out
