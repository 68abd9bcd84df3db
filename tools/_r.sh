timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x 2>&1 | tail -4
timeout 200 python tools/trace_mix.py 2>&1 | tail -24
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ml.json 2> gpurun_out/bench_ml.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_ml.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','ms_per_step_min','e2e','gpu_launches')}); print({k: round(v['mean_us'],1) for k,v in d['kernels'].items()}); print(d['reps_ms_per_step'], d['clocks'])"
