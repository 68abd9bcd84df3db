"""Train-step time (fwd + nmse.backward + FusedAdam.step) on this GPU: cavity B=64 and cylinder B=256, fp32 and bf16 storage."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cfdbench_b200 import synth
for problem, b in (("cavity", 64), ("cylinder", 256)):
    for act in ("f32", "bf16"):
        r = bench.timed_train_step(synth.n_case_params(problem), b, steps=10, warmup=3, problem=problem, act=act)
        print(f"{problem:8s} B={b:3d} {act:4s}: {r['ms_per_step']:.3f} ms/step", flush=True)
