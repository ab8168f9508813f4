#!/bin/bash
# One GPU-box visit that produces everything profiles/ needs for the current HEAD: the whole GPU suite the way the
# driver runs it (one process), smoke, the default bench line, the reference arm, the ncu launch list of the eager
# step, the per-op table, the CUPTI step timeline and one ncu --set full capture of every op of the kernel table.
# usage: bash tools/gpu_evidence.sh TAG [skip-tests]
cd "$(dirname "$0")/.."
TAG=${1:-r2z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_${TAG}.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -q -m gpu --tb=short --no-header -p no:cacheprovider > gpurun_out/gputests_${TAG}.log 2>&1
  echo "gpu tests rc=$? : $(tail -1 gpurun_out/gputests_${TAG}.log)"
  grep -E "^(FAILED|ERROR)" gpurun_out/gputests_${TAG}.log | head -20
fi
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_${TAG}.log
timeout 500 python bench.py --kernels > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref rc=$?"
timeout 200 python tools/time_ops.py > gpurun_out/time_ops_${TAG}.txt 2>&1; echo "time_ops rc=$?"
timeout 200 python tools/step_timeline.py --out gpurun_out/timeline_${TAG}.json > gpurun_out/timeline_${TAG}.txt 2>&1; echo "timeline rc=$?"
timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --step-only > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --step-only > gpurun_out/ncu_${TAG}.log 2>&1; echo "launch list rc=$?"
# one ncu --set full capture of the kernels changed last (regex in $3; the whole table takes > 10 minutes of box time)
KREGEX=${3:-bn_bwd|pw_tc}
timeout 120 python tools/run_kernels.py --only bn_bwd,pwconv_fwd,pwconv_dgrad > gpurun_out/run_kernels_${TAG}.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -o /tmp/prof_${TAG} \
    python tools/run_kernels.py --once --only bn_bwd,pwconv_fwd,pwconv_dgrad > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
cp gpurun_out/run_kernels_ops.json gpurun_out/run_kernels_ops_${TAG}.json
