#!/bin/bash
# One GPU-box visit: GPU tests, a bench run, then (only after the plain run exited 0) the ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1c}
bash tools/gpu_check.sh tests/test_kernels_gpu.py tests/test_model_gpu.py
python bench.py --steps 20 --warmup 5 --kernels > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; cat gpurun_out/bench_${TAG}.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_${TAG}.log 2>&1
echo "ncu rc=$?"
