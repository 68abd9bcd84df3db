"""Host-side mirror of the reference loss interface (reference src/models/loss.py:8-50).

`MseLoss.forward(preds, labels)` returns a dict of 0-dim tensors {mse, rmse, mae[, nmse]};
`get_score_names()` drives `train_auto.evaluate` (reference src/train_auto.py:75,93).
Each value stays on the device and supports `.item()` / `.backward()`.
"""
from __future__ import annotations

from typing import List

import torch
from torch import Tensor, nn


class MseLoss(nn.Module):
    def __init__(self, normalize: bool, is_masked: bool = False):
        super().__init__()
        self.normalize = normalize
        self.is_masked = is_masked

    def get_score_names(self) -> List[str]:
        return ["mse", "rmse", "mae"] + (["nmse"] if self.normalize else [])

    def forward(self, preds: Tensor, labels: Tensor) -> dict:
        diff = preds - labels
        mse = torch.mean(diff * diff)
        out = {"mse": mse, "rmse": torch.sqrt(mse), "mae": torch.mean(torch.abs(diff))}
        if self.normalize:
            out["nmse"] = mse / torch.mean(labels * labels)
        return out


def loss_name_to_fn(name: str, masked: bool = False) -> MseLoss:
    name = name.lower()
    if masked:
        raise NotImplementedError
    if name == "mse":
        return MseLoss(normalize=False)
    if name == "nmse":
        return MseLoss(normalize=True)
    raise NotImplementedError(name)
