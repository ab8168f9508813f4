#!/bin/bash
# Round-end evidence run on the GPU box: full GPU test suite, smoke, bench (with the CPU baseline), the
# ncu launch list of the eager step, and one ncu --set full capture of every op of the kernel table.
cd "$(dirname "$0")/.."
TAG=${1:-final}
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_kernels_gpu.py tests/test_model_gpu.py
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_${TAG}.log
timeout 400 python bench.py --steps 20 --warmup 5 --kernels > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref rc=$?"
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_${TAG}.log 2>&1; echo "launch list rc=$?"
timeout 120 python tools/run_kernels.py > gpurun_out/run_kernels_${TAG}.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dw_|pw_tc|wgrad_tc|bn_|upsample|stem' -o /tmp/prof_${TAG} \
    python tools/run_kernels.py --once > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
cp gpurun_out/run_kernels_ops.json gpurun_out/run_kernels_ops_${TAG}.json
