import os, sys, torch
sys.path.insert(0, os.getcwd())
from bench import build_model
from cfdbench_b200 import synth
for B in (2, 80, 256):
    m, sd = build_model("bf16", 5)
    m.graph_rollout = False
    batch = synth.make_batch(1, B, "cavity", with_label=False)
    inp, cp, mk = (torch.from_numpy(batch[k]).cuda() for k in ("inputs", "case_params", "mask"))
    for it in range(3):
        with torch.no_grad():
            y = m.generate(inp, cp, mk)
        torch.cuda.synchronize()
        print("B", B, "iter", it, "ok", float(y.abs().mean()), flush=True)
