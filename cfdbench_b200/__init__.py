"""cfdbench_b200 -- B200-native (sm_100a) FNO hot path for CFDBench.

Public surface mirrors the reference's for this path:
    Fno2d, SpectralConv2d_fast, FnoBlock   (reference src/models/fno/fno2d.py)
    AutoCfdModel                           (reference src/models/base_model.py)
    MseLoss, loss_name_to_fn               (reference src/models/loss.py)
"""
from .base_model import AutoCfdModel
from .loss import MseLoss, loss_name_to_fn

__all__ = ["AutoCfdModel", "MseLoss", "loss_name_to_fn", "Fno2d", "FnoBlock", "SpectralConv2d_fast", "FusedAdam", "DeviceFrames"]


def __getattr__(name):  # lazy: importing the package must not require the native library
    if name in ("Fno2d", "FnoBlock", "SpectralConv2d_fast"):
        from . import fno2d
        return getattr(fno2d, name)
    if name == "FusedAdam":
        from .optim import FusedAdam
        return FusedAdam
    if name == "DeviceFrames":
        from .data import DeviceFrames
        return DeviceFrames
    raise AttributeError(name)
