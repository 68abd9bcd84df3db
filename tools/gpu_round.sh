#!/bin/bash
# One GPU session: parity tests, bench, ncu launch list + full captures of the top kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [ncu-kernel-regexes...]
TAG=${1:-r01}; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
# launch list of one rollout step (14 launches) after 2 warm-up steps (+ pack kernels: skip 36)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 64 --csv \
   --log-file gpurun_out/launches_$TAG.csv python tools/prof_driver.py --act bf16 --steps 3 > /dev/null 2>&1
echo "ncu launches exit $?"; grep -c gpu__time gpurun_out/launches_$TAG.csv
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 5 -c 1 -f \
     -o gpurun_out/prof_${K}_$TAG python tools/prof_driver.py --act bf16 --steps 2 > gpurun_out/ncu_${K}_$TAG.log 2>&1
  echo "ncu $K exit $?"
done
ls -la gpurun_out | tail -20
