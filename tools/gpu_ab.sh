#!/bin/bash
# Short GPU-box visit: targeted tests of a changed kernel, its per-op times, then step A/B arms (env gates), each arm =
# `bench.py --step-only` (device-timed graph replays of the whole training step, nothing else).
#   gpurun --timeout 600 -- 'bash tools/gpu_ab.sh TAG "test -k expression" "time_ops prefix" "ARM1" "ARM2" ...'   (an arm = "VAR=1 VAR2=3" or "base")
cd "$(dirname "$0")/.."
TAG=${1:-ab}; KEXPR=${2:-batchnorm}; OPS=${3:-bn_}
shift 3
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_fused_paths_gpu.py tests/test_fused_bn_reduction_gpu.py -q -m gpu \
    -k "$KEXPR" --tb=short --no-header -p no:cacheprovider > gpurun_out/abtests_${TAG}.log 2>&1
echo "tests rc=$? : $(tail -1 gpurun_out/abtests_${TAG}.log)"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/abtests_${TAG}.log | head -30
timeout 200 python tools/time_ops.py --only "$OPS" > gpurun_out/abops_${TAG}.txt 2>&1; echo "time_ops rc=$?"; cat gpurun_out/abops_${TAG}.txt | tail -70
for arm in "base" "$@" "base"; do
    name=$(echo "$arm" | tr ' =' '__')
    if [ "$arm" = "base" ]; then envs="TSS_NOOP=1"; else envs="$arm"; fi
    env $envs timeout 200 python bench.py --step-only --steps 40 --warmup 5 > gpurun_out/ab_${TAG}_${name}.json 2> gpurun_out/ab_${TAG}_${name}.err
    echo "$arm rc=$? $(tail -1 gpurun_out/ab_${TAG}_${name}.json | cut -c1-160)"
done
