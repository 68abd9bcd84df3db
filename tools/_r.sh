timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x 2>&1 | tail -4
timeout 200 python tools/trace_fused.py 2>&1 | grep -E "prologue|standalone"
timeout 200 python tools/trace_dft.py 2>&1 | grep -E "prologue|standalone|globaltimer"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_pl.json 2> gpurun_out/bench_pl.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_pl.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','ms_per_step_min','e2e','gpu_launches')}); print({k: round(v['mean_us'],1) for k,v in d['kernels'].items()})"
