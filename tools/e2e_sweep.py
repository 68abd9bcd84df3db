"""e2e (host round trip per step) vs number of batch chunks / CUDA-graph device rollout: tuning helper."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model, timed_e2e, timed_rollout
from cfdbench_b200 import synth
p = synth.n_case_params("cavity")
batch = synth.make_batch(1, 256, "cavity", with_label=False)
for chunks in (1, 2, 4, 8):
    m, _ = build_model("bf16", p)
    m.host_chunks = chunks
    timed_e2e(m, batch, 5, 2)
    t, _, _ = timed_e2e(m, batch, 20, 2)
    print(f"chunks={chunks}: e2e {20 / t:.1f} steps/s ({1e3 * t / 20:.3f} ms/step)")
    del m
inp, cp, mk = (torch.from_numpy(batch[k]).cuda() for k in ("inputs", "case_params", "mask"))
for graph in (False, True):
    m, _ = build_model("bf16", p)
    m.graph_rollout = graph
    t, _ = timed_rollout(m, inp, cp, mk, 20, 3)
    print(f"graph={graph}: device rollout {20 / t:.1f} steps/s")
    del m
