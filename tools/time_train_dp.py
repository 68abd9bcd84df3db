"""torchrun helper: train-step time (cavity B=64/GPU and cylinder B=256/GPU, fp32 storage, FusedAdam) for each way of
all-reducing the gradient buffer.  python -m torch.distributed.run --nproc-per-node N tools/time_train_dp.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cfdbench_b200 import dp, synth
rank, local, world = dp.init_from_env("nccl")
torch.cuda.set_device(local)
out = {}
for mode in ("one", "two", "all"):
    os.environ["FNO_DP_SEGMENTS"] = mode
    a = bench.timed_train_step(5, 64, steps=10, warmup=4)
    b = bench.timed_train_step(8, 256, steps=6, warmup=3, problem="cylinder")
    out[mode] = {"cavity_b64_ms": round(a["ms_per_step"], 3), "cylinder_b256_ms": round(b["ms_per_step"], 3)}
if rank == 0:
    print(json.dumps({"world": world, **out}))
torch.distributed.destroy_process_group() if world > 1 else None
