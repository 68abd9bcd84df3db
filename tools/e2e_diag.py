"""Per-step wall time of the chunked host round trip, several fresh model instances: looks for outliers."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model
from cfdbench_b200 import synth
p = synth.n_case_params("cavity")
batch = synth.make_batch(1, 256, "cavity", with_label=False)
pin = {k: torch.from_numpy(batch[k]).pin_memory() for k in ("inputs", "case_params", "mask")}
for chunks in (2, 4, 2, 4, 2, 4, 1, 8):
    m, _ = build_model("bf16", p)
    m.host_chunks = chunks
    cur = pin["inputs"]
    ts = []
    for i in range(30):
        t0 = time.perf_counter()
        cur = m.generate_many(cur, pin["case_params"], pin["mask"], 1)[0]
        ts.append(1e3 * (time.perf_counter() - t0))
    ts = np.array(ts[5:])
    print(f"chunks={chunks}: mean {ts.mean():.3f} ms  median {np.median(ts):.3f}  min {ts.min():.3f}  max {ts.max():.3f}  "
          f"first5 {[round(x, 2) for x in ts[:5]]}", flush=True)
    del m
