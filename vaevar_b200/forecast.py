"""Python handle on the forecast network LGUnet_all_1 inside libvaevar.so (vv_net1_*, include/vaevar.h).

    init_model_forecast   da_4dvar.py:548-569    LGUnet_all_1(**params) + load_state_dict + eval
    integrate             da_4dvar.py:666-681    (x - mean) / std -> model(.)[:, :69] -> * std + mean, once per cycle (:1329)

Plumbing only: torch supplies device memory and the stream; forward only (the DA loop never differentiates the forecast).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from . import _lib
from .config import Net1Config, era5_stats
from .engine import _dev32, _ptr, _stream


def _net1_c(cfg: Net1Config, keep_out: int) -> _lib.Net1ConfigC:
    if tuple(cfg.patch_size) != (3, 2) or tuple(cfg.stride) != (2, 2):
        raise NotImplementedError("only patch_size (3, 2) / stride (2, 2) (model_0.25degree/training_options.yaml) is built")
    c = _lib.Net1ConfigC()
    c.img_h, c.img_w = cfg.img_size
    c.n_groups = cfg.groups
    for i, v in enumerate(cfg.inchans_list):
        c.in_chans[i] = v
    for i, v in enumerate(cfg.outchans_list):
        c.out_chans[i] = v
    c.enc_dim, c.embed_dim = cfg.enc_dim, cfg.embed_dim
    c.win_h, c.win_w = cfg.window_size
    c.n_levels = len(cfg.enc_depths)
    for i, (d, h) in enumerate(zip(cfg.enc_depths, cfg.enc_heads)):
        c.enc_depth[i], c.enc_heads[i] = d, h
    c.n_lg = len(cfg.lg_depths)
    for i, (d, h) in enumerate(zip(cfg.lg_depths, cfg.lg_heads)):
        c.lg_depth[i], c.lg_heads[i] = d, h
    c.keep_out = keep_out
    return c


class ForecastNet:
    """One LGUnet_all_1 on one B200.  keep_out = 69 keeps the mean channels only, as `model(z)[:, :69]` does (da_4dvar.py:674)."""

    def __init__(self, cfg: Net1Config, keep_out: int = 69, device: str = "cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("vaevar_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        torch.cuda.current_stream()
        _lib.check(self.lib.vv_set_device(self.device.index or 0))
        self.cfg = cfg
        self.n_in = cfg.in_chans
        self.n_out = keep_out or cfg.out_chans
        c = _net1_c(cfg, keep_out)
        self._h = C.c_void_p()
        _lib.check(self.lib.vv_net1_create(C.byref(c), C.byref(self._h)))
        mean, std, _ = era5_stats()
        if self.n_in == len(mean) and self.n_out == len(mean):
            self.set_constants(mean, std)

    def load_state_dict(self, sd: Dict[str, "np.ndarray | torch.Tensor"]):
        """Keys are the reference state_dict names (a "module." prefix is stripped as da_4dvar.py:560-566 does)."""
        for k, v in sd.items():
            if k.startswith("module."):
                k = k[7:]
            t = _dev32(v, self.device)
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self.lib.vv_net1_set_weight(self._h, k.encode(), _ptr(t), shape, t.dim()))
        torch.cuda.synchronize()

    def finalize(self):
        _lib.check(self.lib.vv_net1_finalize(self._h))

    def set_constants(self, mean, std):
        a = [np.ascontiguousarray(np.asarray(x, np.float32)) for x in (mean, std)]
        _lib.check(self.lib.vv_net1_set_constants(self._h, *[x.ctypes.data_as(C.c_void_p) for x in a]))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(sum C_in, H, W) normalised -> (keep_out, H, W); LGUnet_all_1.forward on one sample (networks/LGUnet_all.py:772-777)."""
        x = _dev32(x, self.device)
        assert tuple(x.shape) == (self.n_in, *self.cfg.img_size), tuple(x.shape)
        out = torch.empty(self.n_out, *self.cfg.img_size, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.vv_net1_forward(self._h, _ptr(x), _ptr(out), _stream()))
        return out

    def integrate(self, xa: torch.Tensor, steps: int = 1) -> torch.Tensor:
        """(69, H, W) physical -> physical after `steps` model applications (da_4dvar.py:666-681 with interpolation=False)."""
        xa = _dev32(xa, self.device)
        assert tuple(xa.shape) == (self.n_in, *self.cfg.img_size), tuple(xa.shape)
        out = torch.empty_like(xa)
        _lib.check(self.lib.vv_net1_integrate(self._h, _ptr(xa), _ptr(out), int(steps), _stream()))
        return out

    def profile_ops(self, reps: int = 3):
        """Per-launch steady-state times of one application: list of dicts (kind, ms, flop, shape)."""
        cap = 4096
        ms = (C.c_float * cap)(); kind = (C.c_int * cap)(); fl = (C.c_double * cap)(); mnk = (C.c_int * (4 * cap))()
        n = self.lib.vv_net1_profile_ops(self._h, int(reps), ms, kind, fl, mnk, cap)
        if n < 0:
            _lib.check(n)
        names = {0: "gemm", 1: "ln_fwd", 7: "rope2", 8: "sd_attn", 9: "patch_embed32", 10: "convt_head32"}
        return [{"kind": names.get(kind[i], str(kind[i])), "ms": ms[i], "flop": fl[i], "shape": [mnk[4 * i + j] for j in range(4)]} for i in range(min(n, cap))]

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.vv_net1_last_launch_count(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self.lib.vv_net1_device_bytes(self._h))

    def close(self):
        if self._h:
            self.lib.vv_net1_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
