#!/bin/bash
# GPU-box visit for the kernels that are built but gated off (tests/test_fused_paths_gpu.py): every step runs
# under its own timeout so that an unvalidated kernel cannot hold the box.  Then an A/B of the training step with
# each gate on, same bench command as the default arm (the gate that wins AND passes becomes the default).
#   gpurun --timeout 1500 -- 'bash tools/gpu_experimental.sh r2x'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2x}
# no -x: one visit should report every gate that fails, not only the first
TSS_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_fused_paths_gpu.py -q -rf --timeout 120 > gpurun_out/experimental_${TAG}.log 2>&1
echo "experimental tests rc=$?"; grep -E "^FAILED|passed|failed|error" gpurun_out/experimental_${TAG}.log | tail -40
timeout 600 python tools/experimental_kernels.py > gpurun_out/experimental_kernels_${TAG}.jsonl 2> gpurun_out/experimental_kernels_${TAG}.err
echo "experimental kernels rc=$?"; cat gpurun_out/experimental_kernels_${TAG}.jsonl | cut -c1-220
for arm in "base" "TSS_FUSE_BNRED_EXT=1" "TSS_FUSE_BNAPPLY=1" "TSS_FUSE_PPM=1" "TSS_FUSE_BNAPPLY_DW=1" "TSS_FUSE_BNFIN=1" "TSS_FUSE_BNIN=1" "TSS_FUSE_BNIN_PW=1" "TSS_STEM_TC=1" "TSS_STEM_TC=1 TSS_STEM_BWD_FUSED=1" "TSS_DEFER_LOGITS=1" "TSS_OWN_DROPOUT=1" "TSS_FUSE_BNRED_EXT=1 TSS_FUSE_BNAPPLY=1 TSS_FUSE_BNAPPLY_DW=1 TSS_FUSE_PPM=1 TSS_FUSE_BNFIN=1 TSS_FUSE_BNIN=1 TSS_FUSE_BNIN_PW=1 TSS_STEM_TC=1 TSS_STEM_BWD_FUSED=1 TSS_DEFER_LOGITS=1 TSS_OWN_DROPOUT=1"; do
    name=$(echo "$arm" | tr ' =' '__')
    if [ "$arm" = "base" ]; then envs=""; else envs="$arm"; fi
    env $envs timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${TAG}_${name}.json 2> gpurun_out/ab_${TAG}_${name}.err
    echo "$arm rc=$? $(python -c "import json,sys; d=json.loads(open('gpurun_out/ab_${TAG}_${name}.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['gpu_launches'])" 2>&1 | tail -1)"
done
# confusion matrix over 500 random maps (BASELINE.json configs[3]): default kernel vs the warp-private variant
for v in 0 1; do
    TSS_CM_VARIANT=$v timeout 300 python tools/bench_configs.py --config 4 > gpurun_out/cm_${TAG}_variant$v.json 2> gpurun_out/cm_${TAG}_variant$v.err
    echo "TSS_CM_VARIANT=$v rc=$? $(tail -1 gpurun_out/cm_${TAG}_variant$v.json | cut -c1-300)"
done
# inference batch sweep with the grouped pyramid path and the tensor-core stem
for arm in "base" "TSS_FUSE_PPM=1 TSS_STEM_TC=1"; do
    name=$(echo "$arm" | tr ' =' '__')
    if [ "$arm" = "base" ]; then envs=""; else envs="$arm"; fi
    env $envs timeout 300 python tools/bench_configs.py --config 5 --batches 1,16 > gpurun_out/inf_${TAG}_${name}.json 2> gpurun_out/inf_${TAG}_${name}.err
    echo "inference $arm rc=$? $(tail -1 gpurun_out/inf_${TAG}_${name}.json | cut -c1-300)"
done
# end-to-end leg fed with uint8 frames (device input pipeline in front of the graphed step)
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-uint8 > gpurun_out/ab_${TAG}_e2e_uint8.json 2> gpurun_out/ab_${TAG}_e2e_uint8.err
echo "e2e-uint8 rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/ab_${TAG}_e2e_uint8.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e'])" 2>&1 | tail -1)"
# end-to-end leg with one graph per staging slot (no device-to-device copy of the batch)
TSS_SLOT_GRAPHS=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${TAG}_slot_graphs.json 2> gpurun_out/ab_${TAG}_slot_graphs.err
echo "slot-graphs rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/ab_${TAG}_slot_graphs.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e'])" 2>&1 | tail -1)"
# ContextNet-14 training step (BASELINE.json configs[2], 8 x 1024 x 2048 per GPU): default vs every training gate
ALL="TSS_FUSE_BNRED_EXT=1 TSS_FUSE_BNAPPLY=1 TSS_FUSE_BNAPPLY_DW=1 TSS_FUSE_BNFIN=1 TSS_FUSE_BNIN=1 TSS_FUSE_BNIN_PW=1 TSS_STEM_TC=1 TSS_STEM_BWD_FUSED=1 TSS_DEFER_LOGITS=1 TSS_OWN_DROPOUT=1"
for arm in "base" "$ALL"; do
    if [ "$arm" = "base" ]; then envs=""; name=base; else envs="$arm"; name=all; fi
    env $envs timeout 400 python tools/bench_configs.py --config 3 --steps 10 > gpurun_out/cn_${TAG}_${name}.json 2> gpurun_out/cn_${TAG}_${name}.err
    echo "contextnet14 $name rc=$? $(tail -1 gpurun_out/cn_${TAG}_${name}.json | cut -c1-300)"
done
