"""e2e step time (host frame in -> host prediction out, B=256 bf16) for chunk plans of the staged host path."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cfdbench_b200 import synth
batch = synth.make_batch(1, 256, "cavity", with_label=False)
plans = [None, (0.25, 0.75), (0.375, 0.625), (0.75, 0.25), (0.25, 0.5, 0.25), (0.125, 0.5, 0.375), (0.1875, 0.8125), (0.25, 0.375, 0.375)]
for plan in plans:
    m, _ = bench.build_model("bf16", 5)
    m.host_chunk_plan = plan
    reps = sorted(bench.timed_e2e(m, batch, 20, 3)[0] for _ in range(3))
    print(f"plan {str(plan):24s}: {1e3 * reps[1] / 20:.3f} ms/step  ({20 / reps[1]:.0f} steps/s)", flush=True)
    del m
