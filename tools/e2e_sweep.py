"""e2e step time (host frame in -> host prediction out, B=256 bf16) for the host-path variants."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cfdbench_b200 import synth
p = 5
batch = synth.make_batch(1, 256, "cavity", with_label=False)
for name, zc, zch, hch in (("zero-copy x1", True, 1, 4), ("zero-copy x2", True, 2, 4), ("zero-copy x4", True, 4, 4),
                           ("staged 4 chunks", False, 1, 4), ("staged 2 chunks", False, 1, 2)):
    m, _ = bench.build_model("bf16", p)
    m.host_zero_copy, m.zero_copy_chunks, m.host_chunks = zc, zch, hch
    reps = sorted(bench.timed_e2e(m, batch, 20, 3)[0] for _ in range(3))
    print(f"{name:18s}: {1e3 * reps[1] / 20:.3f} ms/step  ({20 / reps[1]:.0f} steps/s)", flush=True)
    del m
