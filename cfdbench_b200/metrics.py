"""Device-side rollout evaluation.

`multistep_metrics` reproduces what `test_multistep.infer` reports (reference src/test_multistep.py:153-177):
for every rollout step, the mean over cases of `get_metrics(preds_u * mask, label_u * mask)`
(`mse`, `nmse = mse / mean(label^2)`, `mae`; reference :73-83) -- with ONE kernel launch and ONE device->host copy
instead of three `.item()` synchronisations per step and case.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Union

import torch
from torch import Tensor

from . import _lib

HW = 64 * 64


def multistep_metrics(preds: Union[Tensor, Sequence[Tensor]], label_u: Tensor, mask: Tensor) -> List[Dict[str, float]]:
    """preds: (S,B,2,64,64) tensor or the list `generate_many` returns; label_u, mask: (S,B,64,64) CUDA tensors
    (the reference compares step s with frame s of the case, u channel only, using that frame's mask).
    Returns a list of S dicts {mse, nmse, mae}, each the mean over the B cases (as `combine_dicts` does)."""
    if not isinstance(preds, Tensor):
        preds = torch.stack(list(preds))
    if preds.device.type != "cuda":
        raise _lib.FnoNativeError("multistep_metrics has no CPU path: pass CUDA tensors")
    s, b = preds.shape[:2]
    if tuple(preds.shape[2:]) != (2, 64, 64) or tuple(label_u.shape) != (s, b, 64, 64) or tuple(mask.shape) != (s, b, 64, 64):
        raise ValueError("expected preds (S,B,2,64,64), label_u (S,B,64,64), mask (S,B,64,64)")
    preds = preds.contiguous().float()
    label_u = label_u.to(preds.device).contiguous().float()
    mask = mask.to(preds.device).contiguous().float()
    sums = torch.empty(s, b, 3, dtype=torch.float32, device=preds.device)
    lib = _lib.load()
    with torch.cuda.device(preds.device):
        st = C.c_void_p(torch.cuda.current_stream(preds.device).cuda_stream)
        _lib.check(lib.fno_multistep_metrics(preds.data_ptr(), label_u.data_ptr(), mask.data_ptr(), sums.data_ptr(),
                                             s, b, st), "fno_multistep_metrics")
    host = sums.double().cpu()  # the only synchronisation
    mse = host[..., 0] / HW
    nmse = mse / (host[..., 1] / HW)
    mae = host[..., 2] / HW
    return [dict(mse=float(mse[i].mean()), nmse=float(nmse[i].mean()), mae=float(mae[i].mean())) for i in range(s)]
