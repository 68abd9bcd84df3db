"""Per-kernel duration vs batch: separates each kernel's fixed cost (launch, prologue, pipeline fill) from its slope."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model, kernel_pass
from cfdbench_b200 import synth
p = synth.n_case_params("cavity")
res = {}
bs = (32, 64, 128, 256, 512)
for b in bs:
    batch = synth.make_batch(1, b, "cavity", with_label=False)
    inp, cp, mk = (torch.from_numpy(batch[k]).cuda() for k in ("inputs", "case_params", "mask"))
    m, _ = build_model("bf16", p)
    kernel_pass(m, inp, cp, mk, 5)
    res[b] = kernel_pass(m, inp, cp, mk, 20)
    del m
names = list(res[bs[0]].keys())
print("kernel      " + "".join(f"B={b:<8d}" for b in bs) + " fixed(us)  per-128(us)")
for n in names:
    t = np.array([res[b][n]["mean_us"] for b in bs])
    slope, icpt = np.polyfit(np.array(bs[2:]), t[2:], 1)
    print(f"{n:11s} " + "".join(f"{v:<10.1f}" for v in t) + f" {icpt:8.1f} {slope * 128:10.1f}")
