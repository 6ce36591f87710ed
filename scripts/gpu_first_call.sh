#!/bin/bash
# First GPU call of a round: everything that was written without hardware access gets its first run, one process per
# group so that a faulting kernel cannot poison the others.  Logs under gpurun_out/.
#   gpurun --timeout 900 -- bash scripts/gpu_first_call.sh
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ==="; timeout ${LIMIT:-300} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -n ${TAIL:-6} gpurun_out/$name.log; }
# 1. pending tests (xfail markers ignored so that real pass/fail shows)
for f in test_gpu_z_gemm_mma test_gpu_z_audio test_gpu_z_data test_gpu_zz_dx_tc; do
  run $f python -m pytest tests/$f.py -q -m gpu --runxfail --tb=short -p no:cacheprovider
done
# 2. the validated suites must still be green with the rebuilt library
run train python -m pytest tests/test_gpu_train.py -q -m gpu --tb=short -p no:cacheprovider
# 3. where the training step spends its time, with both GEMM kernels
run sections_ffma python scripts/train_sections.py
cp gpurun_out/train_sections.txt gpurun_out/train_sections_ffma.txt 2>/dev/null
TACO_GEMM_IMPL=1 run sections_mma python scripts/train_sections.py
cp gpurun_out/train_sections.txt gpurun_out/train_sections_mma.txt 2>/dev/null
# 4. smoke + the full bench line (train / c5 side measurements included)
run smoke python __graft_entry__.py --smoke
LIMIT=600 TAIL=3 run bench python bench.py --steps 10 --warmup 3
