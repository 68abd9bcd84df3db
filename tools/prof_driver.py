"""Tiny driver for ncu: N rollout steps of the BASELINE configs[1] batch (B=256 cavity) through the C ABI.
Usage: python tools/prof_driver.py [--act bf16|f32] [--steps 3] [--batch 256] [--train]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model  # noqa: E402
from cfdbench_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--act", default="bf16")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--train", action="store_true")
a = ap.parse_args()
p = synth.n_case_params("cavity")
model, _ = build_model(a.act, p)
batch = synth.make_batch(1, a.batch, "cavity", with_label=True)
tb = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
if a.train:
    for _ in range(a.steps):
        out = model(**tb)
        out["loss"]["nmse"].backward()
else:
    model.generate_many(tb["inputs"], tb["case_params"], tb["mask"], a.steps)
torch.cuda.synchronize()
print("done")
