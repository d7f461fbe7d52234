"""The VAE-Var inner loop behind the reference driver's call surface.

    cyclic_4dvar.integrate      da_4dvar.py:666-681
    cyclic_4dvar.one_step_DA    da_4dvar.py:933, vae4dvar branch :1179-1306
    cal_loss / closure / LBFGS  da_4dvar.py:1210-1246, 1298-1299
    Metrics.WRMSE / Bias        utils/metrics.py:282-296, 65-82 (kept bit-for-bit incl. pi ~ 3.1416)

The latent z, its gradient, the L-BFGS history and the trajectory never leave the GPU; per closure evaluation the
controller reads back four doubles.  The optimisation runs on the network grid, where the reference's resampling to 721x1440
(vae.py:90, da_4dvar.py:671,679) is the identity; `integrate(..., interpolation=True)` (the forecast of a native-resolution
analysis) goes through the seam kernels of seams.py.
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import numpy as np
import torch

from .config import NetConfig, era5_stats
from .engine import LBFGS, Engine


def lat_weight(num_lat: int, device, dtype=torch.float32) -> torch.Tensor:
    """utils/metrics.py:5-10 -- note the literal 3.1416."""
    j = torch.arange(0, num_lat, device=device)
    lat = 90.0 - j * 180.0 / float(num_lat - 1)
    cosl = torch.cos(3.1416 / 180.0 * lat)
    return (num_lat * cosl / torch.sum(cosl)).to(dtype).reshape(1, 1, -1, 1)


def wrmse(pred, gt, data_std):
    w = lat_weight(pred.shape[2], pred.device)
    return torch.mean(torch.sqrt(torch.mean(w * (pred - gt) ** 2.0, dim=(-1, -2))), dim=0) * data_std


def bias(pred, gt, data_std):
    w = lat_weight(pred.shape[2], pred.device)
    return torch.mean(torch.mean(w * (pred - gt), dim=(-1, -2)), dim=0) * data_std


class obs_interpolater:
    """da_4dvar.py:62-94: linear interpolation in log-pressure between the 13 model levels and `dim_out` observation levels
    (`interp`, (dim_out, dim_in)) and back (`interp_inv`, (dim_in, dim_out)); float32 tensors like the reference's."""

    def __init__(self, dim_in: int = 13, dim_out: int = 40, device="cpu"):
        self.dim_in, self.dim_out, self.device = dim_in, dim_out, device
        self.height_level = [50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000]
        self.height_level_new = np.round(np.exp(np.linspace(3.91202301, 6.90775528, dim_out)))
        self.interp = self._between(self.height_level_new, np.asarray(self.height_level, np.float64)).to(device)
        self.interp_inv = self._between(np.asarray(self.height_level, np.float64), self.height_level_new).to(device)

    @staticmethod
    def _between(targets, nodes) -> torch.Tensor:
        """Row i: weights of `nodes` that reproduce targets[i]: 1 on an exact hit, otherwise the two log-linear weights of the
        bracketing pair (strictly inside it); rows outside every bracket stay zero, as in the reference's double loop."""
        w = torch.zeros(len(targets), len(nodes))
        ln = np.log(nodes)
        for i, tv in enumerate(targets):
            hit = np.flatnonzero(nodes == tv)
            for j in hit:
                w[i, j] = 1
            j = int(np.searchsorted(nodes, tv)) - 1
            if len(hit) == 0 and 0 <= j < len(nodes) - 1 and nodes[j] < tv < nodes[j + 1]:
                d = ln[j + 1] - ln[j]
                w[i, j] = (ln[j + 1] - np.log(tv)) / d
                w[i, j + 1] = (np.log(tv) - ln[j]) / d
        return w


class VaeVar4D:
    """Engine-backed stand-in for `cyclic_4dvar` restricted to da_mode == "vae4dvar"."""

    def __init__(self, dec_cfg: NetConfig, flow_cfg: Optional[NetConfig], dec_sd: Dict, flow_sd: Optional[Dict],
                 da_win: int = 1, Nit: int = 4, obs_coeff: float = 1.0, device: str = "cuda:0",
                 recompute: bool = False, use_graph: bool = True, verbose: bool = True, engine: Optional[Engine] = None,
                 obs_type: str = "free", interp_dim: int = 40, forecast_model=None):
        self.forecast_model = forecast_model                                 # da_4dvar.py:484: LGUnet_all_1 on the analysis grid, or None
        self.obs_type = obs_type                                             # da_4dvar.py:476; "real..." selects the augmented obs space
        self.obs_interp = obs_interpolater(13, interp_dim)                   # da_4dvar.py:493
        self.da_win, self.Nit, self.obs_coeff, self.verbose = da_win, Nit, obs_coeff, verbose
        self.device = torch.device(device)
        if engine is not None:                       # an already loaded engine (same networks, same window length)
            assert engine.T == da_win, "engine was built for another window length"
            self.engine = engine
        else:
            self.engine = Engine(dec_cfg, flow_cfg, T=da_win, recompute=recompute, use_graph=use_graph, device=device)
            self.engine.load_state_dict(0, dec_sd)
            if flow_cfg is not None:
                self.engine.load_state_dict(1, flow_sd)
            self.engine.finalize()
        mean, std, _ = era5_stats()
        self.model_mean, self.model_std = mean, std                        # float64, da_4dvar.py:645-646
        self.model_mean_gpu = torch.from_numpy(mean).float().to(self.device)
        self.model_std_gpu = torch.from_numpy(std).float().to(self.device)
        self.nchannel = 69
        self.nlat, self.nlon = dec_cfg.img_size
        self.latent = dec_cfg.in_chans
        self.metrics_list = {k: [] for k in ("bg_wrmse", "bg_bias", "ana_wrmse", "ana_bias")}
        self.history = []
        self._opt = None
        self._native = False

    def integrate(self, xa: torch.Tensor, model=None, step: int = 1, interpolation: bool = False, detach: bool = True):
        """(69,nlat,nlon) physical -> physical after `step` applications of the flow model (da_4dvar.py:666-681).  With
        `interpolation` the field is brought to the network grid and back with the reference's nearest rule (:670-671, 678-679);
        the per-channel (de)normalisation commutes with that index map, so it stays inside the engine call."""
        xa = xa.to(self.device, torch.float32)
        if model is not None and hasattr(model, "integrate") and not interpolation:
            # run_assimilation's forecast: the native-resolution LGUnet_all_1 applied to the analysis on its own grid (da_4dvar.py:1329)
            return model.integrate(xa, step)
        if tuple(xa.shape[-2:]) == (self.nlat, self.nlon):
            return self.engine.integrate(xa, step)
        if not interpolation and self.verbose:
            # no forecast model on the analysis grid was given: the flow model behind the two seams stands in for it
            print("integrate: field is on the analysis grid, resampling to the network grid and back", flush=True)
        from .seams import resample_nearest
        x = self.engine.integrate(resample_nearest(xa, (self.nlat, self.nlon)), step)
        return resample_nearest(x, tuple(xa.shape[-2:]))

    def _diagnostics(self, z, gt0):
        """WRMSE / Bias of the current analysis (da_4dvar.py:1256-1264) without leaving the device: the fused metric
        kernel normalises both fields and applies utils/metrics.py's latitude weighting in one pass."""
        return self.engine.metrics(self.engine.decode_native(z) if self._native else self.engine.decode(z), gt0)

    def one_step_DA(self, gt, xb, yo, H, R, mode: str = "vae4dvar"):
        if mode != "vae4dvar":
            raise NotImplementedError("not implemented da mode")            # da_4dvar.py:1308-1309
        dev = self.device
        gt0 = torch.as_tensor(gt[0]).to(dev, torch.float32)
        # fields on a finer grid than the networks' (the reference's 721x1440 over 128x256): decoder_hr / integrate(..., True, False)
        self._native = tuple(torch.as_tensor(xb).shape[-2:]) != (self.nlat, self.nlon)
        if self.obs_type[:4] == "real":                                      # yo / H / R in the 204-channel space, da_4dvar.py:1196-1206
            self._native = True
            self.engine.set_case_real_obs(xb, yo, H, R, self.obs_interp.interp, self.obs_coeff)
        else:
            (self.engine.set_case_native if self._native else self.engine.set_case)(xb, yo, H, R, self.obs_coeff)
        z = torch.zeros(1, self.latent, self.nlat, self.nlon, device=dev)   # da_4dvar.py:1238
        if self._opt is None:                                               # da_4dvar.py:1240: a new optimiser per cycle --
            self._opt = LBFGS(self.engine, history_size=10, max_iter=10)    # here one object, reset (its 30 device vectors are kept)
        opt = self._opt
        opt.reset()
        t0 = time.time()
        for kk in range(self.Nit + 1):
            w, b = self._diagnostics(z, gt0)
            # cal_loss(z), da_4dvar.py:1265: after a step the optimiser already holds the split of the point it stopped on
            J = self.engine.cost(z).cpu() if kk == 0 else opt.last_cost()
            if self.verbose:
                print("iter: %d, RMSE (z500): %.4g Bias (z500): %.4g q500: %.4g, t2m: %.4g t850: %.4g u500: %.4g, v500: %.4g, "
                      "loss reg: %.4g loss obs: %.4g loss: %.4g" % (kk, w[11], b[11], w[24], w[2], w[66], w[37], w[50],
                                                                   J[1], J[2], J[0]), flush=True)
            if kk == 0:
                self.metrics_list["bg_wrmse"].append(w.cpu())
                self.metrics_list["bg_bias"].append(b.cpu())
            elif kk == self.Nit:
                self.metrics_list["ana_wrmse"].append(w.cpu())
                self.metrics_list["ana_bias"].append(b.cpu())
            if kk < self.Nit:
                self.history.append(opt.step(z))                            # lbfgs.step(closure), da_4dvar.py:1298-1299
        xa = self.engine.decode_native(z) if self._native else self.engine.decode(z)
        torch.cuda.current_stream().synchronize()       # this case's stream only: other cases may be in flight on the same GPU
        if self.verbose:
            print("DA finished. Time consumed: %.3f (s)" % (time.time() - t0), flush=True)
        self.z = z
        return xa
