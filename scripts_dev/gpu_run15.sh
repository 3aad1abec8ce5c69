B="python bench.py --workload cfg3 --steps 6 --warmup 3 --no-cpu-baseline --no-side"
$B > gpurun_out/plain_r2b_cfg3_injected.json 2> gpurun_out/plain_r2b_cfg3_injected.err && \
ncu --set full --clock-control none --import-source on -k regex:rollout_injected -s 2 -c 1 -f -o gpurun_out/prof_r2b_cfg3_injected $B > gpurun_out/ncu_r2b_cfg3_injected.log 2>&1
python profiles/summarize_ncu.py gpurun_out/prof_r2b_cfg3_injected.ncu-rep > gpurun_out/r2b_cfg3_injected_ncu_summary.txt 2>&1
rm -f gpurun_out/prof_r2b_cfg3_injected.ncu-rep
head -30 gpurun_out/r2b_cfg3_injected_ncu_summary.txt | grep "duration\|dram__bytes_read.sum  \|inst_executed.sum\|issue_active"
python bench.py --philox-rounds 10 --steps 100 --warmup 10 --no-cpu-baseline --no-injected > gpurun_out/bench_r2b_rounds10.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_r2b_rounds10.json')); print('rounds 10:', d['ms_per_step'], d['roofline']['frac'], 'dense', d['dense_weights']['ms_per_step'], [ (k, round(v['ms_per_step'],4)) for k,v in d['other_configs'].items()])"
