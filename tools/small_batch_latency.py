"""Latency of the reference's actual evaluation pattern (test_multistep.py: B=1, 20 steps per case) with and
without CUDA-graph replay of the native rollout."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model
from cfdbench_b200 import synth
p = synth.n_case_params("cavity")
for b in (1, 8):
    batch = synth.make_batch(1, b, "cavity", with_label=False)
    inp, cp, mk = (torch.from_numpy(batch[k]).cuda() for k in ("inputs", "case_params", "mask"))
    for graph in (False, True):
        m, _ = build_model("f32", p)
        m.graph_rollout = graph
        for _ in range(5):
            m.generate_many(inp, cp, mk, 20)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 50
        for _ in range(n):
            seq = m.generate_many(inp, cp, mk, 20)
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / n
        print(f"B={b} graph={graph}: 20-step rollout {1e3 * t:.3f} ms ({20 / t:.0f} steps/s)")
        del m
