# usage: bash scripts_dev/ab.sh "<bench args>"   -- runs the same bench with _build and _build_alt
for v in _build _build_alt; do
MPPI_B200_LIB=$PWD/mppi_tf_b200/$v/libmppi_b200.so timeout 200 python bench.py $1 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print('$v', d['ms_per_step'], d['e2e']['latency_ms']['p50'], d.get('roofline_injected',{}).get('ms_per_launch'))"
done
