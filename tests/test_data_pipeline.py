"""DeviceFrames reproduces DataLoader + collate_fn (reference src/train_auto.py:33-58, 208-210) from
device-resident frames.  The collate restatement below follows the reference line by line (minus the .cuda() calls)."""
import numpy as np
import pytest
import torch

from cfdbench_b200 import data as cdata


class _FakeAutoDataset(torch.utils.data.Dataset):
    """Same attributes and __getitem__ contract as CavityFlowAutoDataset (reference src/dataset/cavity.py:283-349)."""

    def __init__(self, n=37, n_cases=5, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.inputs = torch.randn(n, 3, 64, 64, generator=g)
        self.inputs[:, 2] = (torch.rand(n, 64, 64, generator=g) > 0.1).float()  # mask channel
        self.labels = torch.randn(n, 3, 64, 64, generator=g)
        self.case_ids = np.sort(np.random.default_rng(seed).integers(0, n_cases, n))
        self.case_params = [dict(density=1.0 + c, viscosity=0.01 * (c + 1), vel_top=0.5 - 0.1 * c, rotated=c % 2,
                                 height=1.0 + 0.25 * c, dx=0.1, width=2.0 - 0.125 * c, dy=0.2) for c in range(n_cases)]

    def __len__(self):
        return len(self.inputs)

    def __getitem__(self, idx):
        return self.inputs[idx], self.labels[idx], self.case_params[self.case_ids[idx]]


def _collate(batch):  # reference src/train_auto.py:33-58 without the .cuda() calls
    inputs, labels, case_params = zip(*batch)
    inputs, labels = torch.stack(inputs), torch.stack(labels)
    labels = labels[:, :-1]
    mask = inputs[:, -1:]
    inputs = inputs[:, :-1]
    keys = [x for x in case_params[0].keys() if x not in ["rotated", "dx", "dy"]]
    vec = [[cp[k] for k in keys] for cp in case_params]
    return dict(inputs=inputs, label=labels, mask=mask, case_params=torch.tensor(vec))


def test_case_table_uses_collate_key_order():
    ds = _FakeAutoDataset()
    t = cdata.case_table(ds.case_params)
    assert t.shape == (5, 5) and t.dtype == np.float32  # density, viscosity, vel_top, height, width
    np.testing.assert_allclose(t[3], [4.0, 0.04, 0.2, 1.75, 1.625], rtol=1e-6)
    assert cdata.EXCLUDED_KEYS == ("rotated", "dx", "dy")


def test_loader_visits_samples_in_dataloader_order():
    """The index stream is built from the samplers DataLoader itself uses, so the same generator gives the same order."""
    ds = _FakeAutoDataset()
    seen = []

    class _Probe(cdata.DeviceFrames):
        def __init__(self, n):  # no GPU: only the index plumbing is under test
            self.n = n

        def batch(self, idx):
            seen.append(list(idx))
            return {}

    for shuffle in (False, True):
        seen.clear()
        list(_Probe(len(ds)).loader(8, shuffle=shuffle, generator=torch.Generator().manual_seed(7)))
        dl = torch.utils.data.DataLoader(range(len(ds)), batch_size=8, shuffle=shuffle,
                                         generator=torch.Generator().manual_seed(7))
        assert seen == [b.tolist() for b in dl]


@pytest.mark.gpu
@pytest.mark.parametrize("frame_dtype", [torch.float32, torch.bfloat16])
def test_device_batches_equal_collate_fn(frame_dtype):
    from cfdbench_b200 import DeviceFrames
    ds = _FakeAutoDataset(n=53, n_cases=6, seed=3)
    frames = DeviceFrames(ds, device="cuda", frame_dtype=frame_dtype)
    assert len(frames) == 53
    g1, g2 = torch.Generator().manual_seed(11), torch.Generator().manual_seed(11)
    dl = torch.utils.data.DataLoader(ds, batch_size=16, shuffle=True, generator=g1, collate_fn=_collate)
    n_batches = 0
    for got, ref in zip(frames.loader(16, shuffle=True, generator=g2), dl):
        n_batches += 1
        for k in ("inputs", "label", "mask", "case_params"):
            r = ref[k].float()
            if frame_dtype == torch.bfloat16 and k != "case_params":
                r = r.to(torch.bfloat16).float()
            assert got[k].device.type == "cuda" and got[k].dtype == torch.float32 and got[k].is_contiguous()
            assert tuple(got[k].shape) == tuple(r.shape), k
            assert torch.equal(got[k].cpu(), r), k
    assert n_batches == 4  # 16 + 16 + 16 + 5: the ragged last batch is kept, as DataLoader(drop_last=False) does
    with pytest.raises(IndexError):
        frames.batch([0, 53])
    # the batch dict feeds the model directly
    from cfdbench_b200 import Fno2d, synth
    from cfdbench_b200.loss import loss_name_to_fn
    m = Fno2d(in_chan=2, out_chan=2, n_case_params=frames.n_case_params, loss_fn=loss_name_to_fn("nmse"), num_layers=4,
              hidden_dim=32, modes1=12, modes2=12).cuda()
    out = m(**frames.batch([1, 2, 3]))
    assert tuple(out["preds"].shape) == (3, 2, 64, 64) and torch.isfinite(out["loss"]["nmse"])
