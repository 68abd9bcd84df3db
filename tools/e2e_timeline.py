"""Where one e2e step (host frame in -> host prediction out, B=256 bf16, staged chunks) spends its time: CUDA events on
the three streams + host clock around the public call, for host_chunks in (1, 2, 4)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cfdbench_b200 import synth
batch = synth.make_batch(1, 256, "cavity", with_label=False)
pin = {k: torch.from_numpy(batch[k]).pin_memory() for k in ("inputs", "case_params", "mask")}
for hch in (1, 2, 4, 8):
    m, _ = bench.build_model("bf16", 5)
    m.host_chunks = hch
    cur = pin["inputs"]
    for _ in range(5):
        cur = m.generate_many(cur, pin["case_params"], pin["mask"], 1)[0]
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        cur = m.generate_many(cur, pin["case_params"], pin["mask"], 1)[0]
        ts.append(time.perf_counter() - t0)
    print(f"host_chunks={hch}: median {1e3 * np.median(ts):.3f} ms/step, min {1e3 * min(ts):.3f}", flush=True)
    if hch >= 2:
        ent = [v for k, v in m._ws_cache.items() if isinstance(k, tuple) and k[0] == "host_chunked"][0]
        s_in, s_cmp, s_out = ent["streams"]
        n = hch
        cb = 256 // n
        # replay the step by hand with timing events
        ev = lambda: torch.cuda.Event(enable_timing=True)
        e_in0, e_in1 = [ev() for _ in range(n)], [ev() for _ in range(n)]
        e_c0, e_c1 = [ev() for _ in range(n)], [ev() for _ in range(n)]
        e_o0, e_o1 = [ev() for _ in range(n)], [ev() for _ in range(n)]
        out = torch.empty(256, 2, 64, 64, pin_memory=True)
        inp = cur
        torch.cuda.synchronize()
        base = ev(); base.record(s_in)
        h0 = time.perf_counter()
        for c in range(n):
            with torch.cuda.stream(s_in):
                e_in0[c].record(s_in)
                ent["d_in"][c].copy_(inp[c * cb:(c + 1) * cb], non_blocking=True)
                e_in1[c].record(s_in)
        h1 = time.perf_counter()
        for c in range(n):
            s_cmp.wait_event(e_in1[c])
            with torch.cuda.stream(s_cmp):
                e_c0[c].record(s_cmp)
                ent["graphs"][c].replay()
                e_c1[c].record(s_cmp)
        h2 = time.perf_counter()
        for c in range(n):
            s_out.wait_event(e_c1[c])
            with torch.cuda.stream(s_out):
                e_o0[c].record(s_out)
                out[c * cb:(c + 1) * cb].copy_(ent["d_out"][c], non_blocking=True)
                e_o1[c].record(s_out)
        h3 = time.perf_counter()
        s_out.synchronize()
        h4 = time.perf_counter()
        f = lambda e: base.elapsed_time(e) * 1e3
        print(f"  host: issue copies {1e6*(h1-h0):.0f} us, issue graphs {1e6*(h2-h1):.0f} us, issue d2h {1e6*(h3-h2):.0f} us, wait {1e6*(h4-h3):.0f} us, total {1e6*(h4-h0):.0f} us")
        for c in range(n):
            print(f"  chunk {c}: h2d {f(e_in0[c]):7.0f} -> {f(e_in1[c]):7.0f}   compute {f(e_c0[c]):7.0f} -> {f(e_c1[c]):7.0f}   d2h {f(e_o0[c]):7.0f} -> {f(e_o1[c]):7.0f}  (us since first event)")
    del m
