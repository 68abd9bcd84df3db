"""Data-parallel training step on 2 GPUs (NCCL): the flat-gradient all-reduce inside backward must give every rank
the mean of the per-shard gradients.  Skipped when fewer than 2 CUDA devices are visible."""
import os

import numpy as np
import pytest
import torch

from cfdbench_b200 import synth

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    from cfdbench_b200 import Fno2d, dp, loss_name_to_fn
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dp.init_from_env("nccl")
    torch.cuda.set_device(rank)
    p = 5
    sd = synth.make_state_dict(41, n_params=p, spectral_gain=50.0)

    def model():
        m = Fno2d(in_chan=2, out_chan=2, n_case_params=p, loss_fn=loss_name_to_fn("nmse"), num_layers=4,
                  hidden_dim=32, modes1=12, modes2=12)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        return m

    batch = synth.make_batch(42, 8, "cavity")
    lo, hi = dp.shard_range(8, rank, world)
    shard = {k: torch.from_numpy(v[lo:hi]).cuda() for k, v in batch.items()}
    m = model()
    m.enable_data_parallel()
    m(**shard)["loss"]["nmse"].backward()
    got = {k: v.grad.detach().clone() for k, v in m.named_parameters()}
    # reference: the same two shards on this GPU without the collective, averaged
    ref = None
    for r in range(world):
        a, b = dp.shard_range(8, r, world)
        m2 = model()
        m2(**{k: torch.from_numpy(v[a:b]).cuda() for k, v in batch.items()})["loss"]["nmse"].backward()
        g = {k: v.grad.detach().clone() for k, v in m2.named_parameters()}
        ref = g if ref is None else {k: ref[k] + g[k] for k in g}
    for k in got:
        e = (got[k] - ref[k] / world).abs().max().item() / (ref[k].abs().max().item() / world + 1e-30)
        assert e < 1e-5, (k, e)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp_gradients_are_the_mean_of_shard_gradients(tmp_path):
    import torch.multiprocessing as mp
    port = 29700 + os.getpid() % 1000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
