for v in _build _build_alt; do
MPPI_B200_LIB=$PWD/mppi_tf_b200/$v/libmppi_b200.so timeout 120 python bench.py --workload cfg4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print('$v', d['ms_per_step'], d['roofline']['frac'])"
done
