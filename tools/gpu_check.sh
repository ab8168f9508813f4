#!/bin/bash
# Runs on the GPU box (through gpurun): GPU test groups in separate processes (a faulting kernel
# must not take the other groups down), logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
make -j8 > gpurun_out/make.log 2>&1 || { echo "BUILD FAILED"; tail -20 gpurun_out/make.log; exit 1; }
status=0
for group in "$@"; do
  name=$(echo "$group" | tr '/:[] ' '_____')
  timeout 600 python -m pytest "$group" -q -m gpu --tb=short --no-header -p no:cacheprovider > "gpurun_out/test_${name}.log" 2>&1
  rc=$?
  echo "== $group -> rc=$rc : $(tail -1 gpurun_out/test_${name}.log)"
  if [ $rc -ne 0 ]; then status=1; grep -E "^(FAILED|ERROR)|Error|error:|assert " "gpurun_out/test_${name}.log" | head -12; fi
done
exit $status
