"""CPU restatement (numpy) of the reference's native-resolution seams.  TEST INFRASTRUCTURE ONLY: imported by tests/, never by the
product path.

    F.interpolate(z, (128, 256)) / F.interpolate(z, (721, 1440)), default mode "nearest"    da_4dvar.py:670-671, 678-679; vae.py:90
    (xa - mean) / std before, z * std + mean after                                          da_4dvar.py:667, 681
    loss_obs = sum(H * (x_pred - yo) ** 2 / R) / 2                                          da_4dvar.py:1207

The index rule lives in a third-party dependency that is not under /root/reference (PyTorch 2.11, ATen UpSample.h
`nearest_idx` / `nearest_neighbor_compute_source_index`; the reference pins no version): identity for equal sizes, dst >> 1 for
an exact doubling, otherwise min(int(floorf(dst * (float(in) / float(out)))), in - 1).  Pinned by tests/golden/seams.npz, which
tools/make_golden_seams.py generates by running torch's own F.interpolate and autograd on CPU (index tables at the reference's
four size pairs, small forward / backward cases with the normalisation on both sides).
"""
from __future__ import annotations

import numpy as np


def nearest_index(out_size: int, in_size: int) -> np.ndarray:
    if out_size == in_size:
        return np.arange(out_size, dtype=np.int64)
    if out_size == 2 * in_size:
        return np.arange(out_size, dtype=np.int64) >> 1
    scale = np.float32(in_size) / np.float32(out_size)
    return np.minimum(np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64), in_size - 1)


def resample(x: np.ndarray, size, mode: int = 0, mean=None, std=None) -> np.ndarray:
    """x (C,Hi,Wi) float32 -> (C,Ho,Wo).  mode 1: normalise, then resample (da_4dvar.py:667, 671); mode 2: resample, then
    de-normalise (:679, 681).  float32 arithmetic with one rounding per operation, like the eager reference."""
    x = np.asarray(x, np.float32)
    ri, ci = nearest_index(size[0], x.shape[1]), nearest_index(size[1], x.shape[2])
    if mode == 1:
        x = (x - np.asarray(mean, np.float32).reshape(-1, 1, 1)) / np.asarray(std, np.float32).reshape(-1, 1, 1)
    y = x[:, ri][:, :, ci]
    if mode == 2:
        y = y * np.asarray(std, np.float32).reshape(-1, 1, 1) + np.asarray(mean, np.float32).reshape(-1, 1, 1)
    return np.ascontiguousarray(y, np.float32)


def resample_adjoint(dout: np.ndarray, in_size, mode: int = 0, std=None) -> np.ndarray:
    """What autograd returns for the input of `resample`: scatter-add in ascending (output row, output column) order in float32 -
    the accumulation order of ATen's CPU backward, which makes the float sums reproducible bit for bit."""
    dout = np.asarray(dout, np.float32)
    c, ho, wo = dout.shape
    ri, ci = nearest_index(ho, in_size[0]), nearest_index(wo, in_size[1])
    if mode == 2:
        dout = dout * np.asarray(std, np.float32).reshape(-1, 1, 1)
    din = np.zeros((c, in_size[0], in_size[1]), np.float32)
    for i in range(ho):
        np.add.at(din[:, ri[i], :], (slice(None), ci), dout[:, i, :])     # unbuffered: columns are added in ascending order
    if mode == 1:
        din = din / np.asarray(std, np.float32).reshape(-1, 1, 1)
    return din


def obs_term(x: np.ndarray, H: np.ndarray, yo: np.ndarray, R: np.ndarray, obs_coeff: float = 1.0):
    """(obs_coeff * loss_obs, its gradient with respect to x) in float64 (da_4dvar.py:1207-1208)."""
    x, H, yo, R = (np.asarray(a, np.float64) for a in (x, H, yo, R))
    r = x - yo
    return obs_coeff * 0.5 * float(np.sum(H * r * r / R)), obs_coeff * H * r / R
