for g in "4,4,5" "3,4,5" "4,3,5" "3,5,5" "2,4,5" "5,3,5" "3,4,4"; do
MPPI_INJ_GEOM=$g timeout 100 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$g', d['roofline_injected']['ms_per_launch'], d['roofline_injected']['frac'])"
done
