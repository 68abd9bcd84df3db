"""Repeat bench.py's timed device rollout in one process and print every call: hunts the intermittent slow state."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model, timed_rollout
from cfdbench_b200 import synth
p = synth.n_case_params("cavity")
batch = synth.make_batch(1, 256, "cavity", with_label=False)
inp, cp, mk = (torch.from_numpy(batch[k]).cuda() for k in ("inputs", "case_params", "mask"))
for rep in range(3):
    m, _ = build_model("bf16", p)
    ts = [timed_rollout(m, inp, cp, mk, 20, 3)[0] for _ in range(3)]
    # direct calls, events around each
    per = []
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(); m.generate_many(inp, cp, mk, 20); e1.record(); torch.cuda.synchronize()
        per.append((round(e0.elapsed_time(e1) / 20 * 1e3), round((time.perf_counter() - w0) / 20 * 1e6)))
    print(f"rep {rep}: timed_rollout us/step {[round(1e6 * t / 20) for t in ts]}  direct (event, wall) {per}", flush=True)
    del m
