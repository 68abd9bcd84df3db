"""Device-resident input pipeline.

The reference holds every (input, label) frame pair of a split in two host tensors `dataset.inputs` /
`dataset.labels` of shape (N, 3, h, w) (u, v, mask), a per-sample `dataset.case_ids` and a list of per-case parameter
dicts `dataset.case_params` (src/dataset/cavity.py:283-331, cylinder.py likewise); `DataLoader` + `collate_fn`
(src/train_auto.py:33-58, 208-210) then builds each batch on the host and copies it to the GPU.  `DeviceFrames`
uploads the split once (fp32, or bf16 to halve its footprint) and produces the same batch dict with one kernel launch
(`fno_gather_batch`); the on-disk format and the dataset classes are untouched -- it takes the dataset object as is.

    frames = DeviceFrames(train_data, device="cuda")
    for batch in frames.loader(batch_size=32, shuffle=True, generator=g):   # same index order as the DataLoader
        out = model(**batch)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Iterator, List, Sequence

import numpy as np
import torch
from torch import Tensor

EXCLUDED_KEYS = ("rotated", "dx", "dy")  # collate_fn, reference src/train_auto.py:45-47


def case_table(case_params: Sequence[dict]) -> np.ndarray:
    """(n_cases, p) float32 table with collate_fn's key order: the keys of the first dict minus EXCLUDED_KEYS."""
    keys = [k for k in case_params[0].keys() if k not in EXCLUDED_KEYS]
    return np.asarray([[cp[k] for k in keys] for cp in case_params], dtype=np.float32).reshape(len(case_params), len(keys))


class DeviceFrames:
    def __init__(self, dataset, device="cuda", frame_dtype: torch.dtype = torch.float32):
        if frame_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("frame_dtype must be float32 or bfloat16")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ValueError("DeviceFrames keeps the split in GPU memory: pass a CUDA device")
        ins, labs = dataset.inputs, dataset.labels
        if ins.dim() != 4 or ins.shape[1] != 3 or tuple(ins.shape[2:]) != (64, 64) or labs.shape != ins.shape:
            raise ValueError(f"expected (N, 3, 64, 64) input / label frames, got {tuple(ins.shape)} / {tuple(labs.shape)}")
        self.device, self.frame_dtype = dev, frame_dtype
        self.n = ins.shape[0]
        self.frames_in = ins.to(device=dev, dtype=frame_dtype).contiguous()
        self.frames_out = labs.to(device=dev, dtype=frame_dtype).contiguous()
        table = case_table(dataset.case_params)
        self.n_case_params = table.shape[1]
        self.case_table = torch.from_numpy(table).to(dev)
        self.case_ids = torch.as_tensor(np.asarray(dataset.case_ids), dtype=torch.int32, device=dev)
        if self.case_ids.numel() != self.n:
            raise ValueError("dataset.case_ids must have one entry per sample")

    def __len__(self) -> int:
        return self.n

    def batch(self, idx) -> Dict[str, Tensor]:
        """The dict collate_fn returns for samples `idx` (reference src/train_auto.py:53-58), all on the device."""
        from . import _lib
        lib = _lib.load()
        idx = torch.as_tensor(idx, dtype=torch.int64)
        if idx.dim() != 1 or idx.numel() == 0:
            raise ValueError("idx must be a non-empty 1-D index list")
        if int(idx.min()) < 0 or int(idx.max()) >= self.n:
            raise IndexError("sample index out of range")
        idx = idx.to(self.device, non_blocking=True)
        b, p, dev = idx.numel(), self.n_case_params, self.device
        out = dict(inputs=torch.empty(b, 2, 64, 64, device=dev), label=torch.empty(b, 2, 64, 64, device=dev),
                   mask=torch.empty(b, 1, 64, 64, device=dev), case_params=torch.empty(b, p, device=dev))
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.fno_gather_batch(self.frames_in.data_ptr(), self.frames_out.data_ptr(),
                                            self.case_table.data_ptr(), self.case_ids.data_ptr(), idx.data_ptr(), b, p,
                                            _lib.ACT_BF16 if self.frame_dtype == torch.bfloat16 else _lib.ACT_F32,
                                            out["inputs"].data_ptr(), out["label"].data_ptr(), out["mask"].data_ptr(),
                                            out["case_params"].data_ptr(), st), "fno_gather_batch")
        idx.record_stream(torch.cuda.current_stream(dev))
        return out

    def batches(self, index_batches: Iterable[Sequence[int]]) -> Iterator[Dict[str, Tensor]]:
        for ib in index_batches:
            yield self.batch(ib)

    def loader(self, batch_size: int, shuffle: bool = False, generator=None, drop_last: bool = False):
        """Batches in exactly the order `DataLoader(dataset, batch_size, shuffle, generator=generator)` visits them:
        the index stream comes from the same torch samplers the DataLoader builds."""
        from torch.utils.data import BatchSampler, RandomSampler, SequentialSampler
        base: List[int] = list(range(self.n))
        sampler = RandomSampler(base, generator=generator) if shuffle else SequentialSampler(base)

        def gen():
            # a DataLoader iterator draws its worker base seed from the generator before the sampler draws the
            # permutation (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__); do the same so that the RNG
            # stream, and with it the visiting order, is identical
            torch.empty((), dtype=torch.int64).random_(generator=generator)
            yield from self.batches(BatchSampler(sampler, batch_size, drop_last))
        return gen()
