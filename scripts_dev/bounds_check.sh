#!/bin/bash
# The safety net that stands in for compute-sanitizer (closed on this pool): the bounds-checked build (make debug:
# -DMPPI_DEBUG_BOUNDS, every computed shared / global index of the rollout kernels asserted) under the ragged-size and
# path-coverage tests.  A violated check traps the kernel and the test fails with a launch error.
# Build here (make -C mppi_tf_b200/csrc debug), run on the GPU box:  gpurun -- 'bash scripts_dev/bounds_check.sh'
mkdir -p gpurun_out
export MPPI_B200_LIB=$PWD/mppi_tf_b200/_build_dbg/libmppi_b200.so
ls -la $MPPI_B200_LIB || exit 1
python -m pytest tests/test_parity_gpu.py tests/test_philox_gpu.py tests/test_philox_variants_gpu.py tests/test_twin_extras_gpu.py \
    tests/test_clip_savgol.py tests/test_kats_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/bounds_check_r2.log 2>&1
echo "bounds-checked build: pytest rc=$?" | tee -a gpurun_out/bounds_check_r2.log
tail -3 gpurun_out/bounds_check_r2.log
