import numpy as np
import torch


def relerr(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (b = reference)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def to_t(batch: dict, device=None, dtype=None):
    out = {}
    for k, v in batch.items():
        if isinstance(v, np.ndarray):
            v = torch.from_numpy(v)
        if torch.is_tensor(v):
            if dtype is not None and v.is_floating_point():
                v = v.to(dtype)
            if device is not None:
                v = v.to(device)
        out[k] = v
    return out


FP32_TOL = 1e-5      # north_star: layer outputs / gradients within 1e-5 relative (fp32 mode)
