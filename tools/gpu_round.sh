#!/bin/bash
# One GPU session: parity tests, bench, ncu launch list + full captures of named kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [kernel_regex:skip ...]
TAG=${1:-r02}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','ms_per_step_min','e2e','gpu_launches')}); print({k: round(v['mean_us'],1) for k,v in d['kernels'].items()}); print(d['roofline']); print(d.get('cpu_baseline'))"; tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
   --log-file gpurun_out/launches_$TAG.csv python tools/prof_driver.py --act bf16 --steps 3 > /dev/null 2>&1
echo "ncu launches exit $?"; grep -c gpu__time gpurun_out/launches_$TAG.csv
for KS in "$@"; do
  K=${KS%%:*}; S=${KS##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f \
     -o gpurun_out/prof_${K}_$TAG python tools/prof_driver.py --act bf16 --steps 2 > gpurun_out/ncu_${K}_$TAG.log 2>&1
  echo "ncu $K exit $?"
done
ls -la gpurun_out | tail -8
