"""Native-resolution seams of the reference on the device, through the C-ABI (include/vaevar.h):

    F.interpolate(x, (721, 1440)) / F.interpolate(z, (128, 256))   nf_model/vae.py:90, da_4dvar.py:671, 679
    the (de)normalisation next to them                              da_4dvar.py:667, 681
    the observation term on the analysis grid                       da_4dvar.py:1207

`interpolate_nearest` is differentiable (torch.autograd.Function over vv_resample_nearest / vv_resample_nearest_adjoint), so
`VAE_lr.decoder_hr` keeps its gradient with respect to z.  No CPU or PyTorch fallback: tensors must live on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from .engine import _ptr, _stream

MODE_PLAIN, MODE_NORMALISE, MODE_DENORMALISE = 0, 1, 2


def _chw(x: torch.Tensor) -> Tuple[int, int, int]:
    if x.dim() < 3:
        raise ValueError("expected a (..., H, W) field")
    if not x.is_cuda or x.dtype != torch.float32:
        raise ValueError("the seam kernels take float32 CUDA tensors (there is no CPU path)")
    h, w = x.shape[-2:]
    return x.numel() // (h * w), h, w


def resample_nearest(x: torch.Tensor, size: Tuple[int, int], mode: int = MODE_PLAIN, mean: Optional[torch.Tensor] = None,
                     std: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(..., Hi, Wi) -> (..., Ho, Wo) with F.interpolate's default (nearest) index rule.  With mode 1 / 2 the leading dimensions
    must flatten to the channel count of `mean` / `std`: mode 1 = (x - mean) / std before the resampling (da_4dvar.py:667, 671),
    mode 2 = * std + mean after it (:679, 681)."""
    x = x.contiguous()
    c, hi, wi = _chw(x)
    ho, wo = int(size[0]), int(size[1])
    if mode != MODE_PLAIN and (mean is None or std is None or mean.numel() != c or std.numel() != c):
        raise ValueError("mode 1 / 2 need one mean and one std per channel")
    out = torch.empty(*x.shape[:-2], ho, wo, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().vv_resample_nearest(_ptr(x), _ptr(out), c, hi, wi, ho, wo, mode,
                                                   _ptr(mean.contiguous()) if mean is not None else None,
                                                   _ptr(std.contiguous()) if std is not None else None, _stream()))
    return out


def resample_nearest_adjoint(dout: torch.Tensor, in_size: Tuple[int, int], mode: int = MODE_PLAIN,
                             std: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Vector-Jacobian product of resample_nearest: (..., Ho, Wo) cotangent -> (..., Hi, Wi); ordered, deterministic sums."""
    dout = dout.contiguous()
    c, ho, wo = _chw(dout)
    hi, wi = int(in_size[0]), int(in_size[1])
    if mode != MODE_PLAIN and (std is None or std.numel() != c):
        raise ValueError("mode 1 / 2 need one std per channel")
    din = torch.empty(*dout.shape[:-2], hi, wi, dtype=torch.float32, device=dout.device)
    with torch.cuda.device(dout.device):
        _lib.check(_lib.load().vv_resample_nearest_adjoint(_ptr(dout), _ptr(din), c, hi, wi, ho, wo, mode,
                                                           _ptr(std.contiguous()) if std is not None else None, _stream()))
    return din


class _Nearest(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size):
        ctx.in_size = tuple(x.shape[-2:])
        return resample_nearest(x, size)

    @staticmethod
    def backward(ctx, dout):
        return resample_nearest_adjoint(dout, ctx.in_size), None


def interpolate_nearest(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """Differentiable stand-in for F.interpolate(x, size) (default mode "nearest") as the reference calls it."""
    return _Nearest.apply(x, tuple(size))


def obs_term(x: torch.Tensor, idx: torch.Tensor, y: torch.Tensor, rinv: torch.Tensor, obs_coeff: float = 1.0, want_grad: bool = True):
    """obs_coeff * 1/2 * sum rinv (x[idx] - y)^2 on a physical-unit field of any size (da_4dvar.py:1207 with H compacted by
    engine.compact_mask); returns (J as a 1-element float64 device tensor, d J / d x or None)."""
    x = x.contiguous()
    if not x.is_cuda or x.dtype != torch.float32:
        raise ValueError("obs_term takes a float32 CUDA field")
    lib = _lib.load()
    J = torch.empty(1, dtype=torch.float64, device=x.device)
    work = torch.empty(int(lib.vv_obs_term_work_doubles()), dtype=torch.float64, device=x.device)
    grad = torch.empty_like(x) if want_grad else None
    with torch.cuda.device(x.device):
        _lib.check(lib.vv_obs_term(_ptr(x), _ptr(idx), _ptr(y), _ptr(rinv), int(idx.numel()), float(obs_coeff), _ptr(J), _ptr(grad),
                                   x.numel(), _ptr(work), _stream()))
    return J, grad
