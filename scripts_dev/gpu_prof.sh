#!/bin/bash
# usage: gpu_prof.sh NAME KERNEL_REGEX SKIP -- bench args...   (plain run first, then one ncu --set full capture; the report is
# summarised on the box and only the text / CSV come back: three full reports exceed gpurun's 64 MiB return limit)
NAME=$1; KREGEX=$2; SKIP=$3; shift 4
B="python bench.py $* --steps 6 --warmup 3 --no-cpu-baseline --no-injected"
mkdir -p gpurun_out
$B > gpurun_out/plain_$NAME.json 2> gpurun_out/plain_$NAME.err && \
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c 1 -f -o gpurun_out/prof_$NAME $B > gpurun_out/ncu_$NAME.log 2>&1
echo "$NAME ncu rc=$?"
if [ -f gpurun_out/prof_$NAME.ncu-rep ]; then
  python profiles/summarize_ncu.py gpurun_out/prof_$NAME.ncu-rep > gpurun_out/${NAME}_ncu_summary.txt 2>&1
  ncu -i gpurun_out/prof_$NAME.ncu-rep --page source --csv > gpurun_out/${NAME}_source.csv 2>/dev/null
  ls -la gpurun_out/prof_$NAME.ncu-rep
  rm -f gpurun_out/prof_$NAME.ncu-rep
fi
