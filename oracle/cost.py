"""ORACLE (test infrastructure, never on the product path).

CPU fp32 restatement of the VAE-Var inner loop of the reference driver, which
itself cannot be imported (da_4dvar.py:18,23-25 pull in petrel_client /
torch_harmonics / xspharm):
    loss(z)            da_4dvar.py:1183-1208   (incl. the real-observation branch :1196-1206 when the
                       Case carries the 40 x 13 level-interpolation matrix of obs_interpolater, :62-82)
    closure()          da_4dvar.py:1242-1246   (autograd supplies the gradient)
    integrate()        da_4dvar.py:666-681     (nlat,nlon parametrised; at the
                       128x256 benchmark grid the nearest resamples at :671,:679
                       and vae.py:90 are identities and are dropped; a Case built
                       with `lr=(h,w)` keeps them: fields on the analysis grid,
                       networks on the (h,w) grid, the reference's real geometry)
    outer L-BFGS loop  da_4dvar.py:1238-1240,1255-1306 with torch.optim.LBFGS as-is
    WRMSE / Bias       utils/metrics.py:282-296, 65-82, 473-474, 526-544
The network applications are oracle.lgunet.lgunet_forward.
"parity unpinned" by the reference's own tests (it has none);
pinned instead against the reference modules run in this container
(tools/make_golden.py -> tests/golden/).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from vaevar_b200.config import NetConfig, era5_stats
from .lgunet import lgunet_forward

Tensor = torch.Tensor


class Case:
    """Tensors captured by the reference closure (da_4dvar.py:1248-1251 and :640-647, :1181)."""

    def __init__(self, case: Dict[str, np.ndarray], obs_coeff: float = 1.0, lr=None, interp=None):
        self.interp = None if interp is None else torch.as_tensor(interp, dtype=torch.float32)    # obs_interp.interp (dim_out, 13)
        self.lr = None if lr is None else tuple(lr)    # network grid when it differs from the analysis grid ((128, 256) in the reference)
        mean, std, stdtr = era5_stats()
        self.mean = torch.from_numpy(mean).float().reshape(-1, 1, 1)     # model_mean_gpu
        self.std = torch.from_numpy(std).float().reshape(-1, 1, 1)       # model_std_gpu
        self.stdTr = torch.tensor(stdtr, dtype=torch.float32).reshape(1, 69, 1, 1)
        self.std64 = torch.from_numpy(std)                               # data_std for the metrics (float64)
        self.xb = torch.from_numpy(case["xb"])
        self.yo = torch.from_numpy(case["yo"])
        self.H = torch.from_numpy(case["H"])
        self.R = torch.from_numpy(case["R"])
        self.gt = torch.from_numpy(case["gt"])
        self.T = self.yo.shape[0]
        self.obs_coeff = obs_coeff

    def to(self, device):
        """Move every captured tensor (bench.py's informational eager-GPU leg runs this same restatement on the B200)."""
        for k, v in list(vars(self).items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self


class OracleNets:
    """The two network applications of the loop as callables: `decode` = VAE_lr.decoder
    (nf_model/vae.py:83-85), `flow` = self.flow_model (da_4dvar.py:571-588).
    tools/make_golden.py substitutes the reference's own nn.Modules behind the same two names."""

    def __init__(self, sd_dec, cfg_dec: NetConfig, sd_flow=None, cfg_flow: Optional[NetConfig] = None):
        self.sd_dec, self.cfg_dec, self.sd_flow, self.cfg_flow = sd_dec, cfg_dec, sd_flow, cfg_flow

    def decode(self, z: Tensor) -> Tensor:
        return lgunet_forward(z, self.sd_dec, self.cfg_dec)

    def flow(self, x: Tensor) -> Tensor:
        return lgunet_forward(x, self.sd_flow, self.cfg_flow)


def integrate(x: Tensor, c: Case, nets, steps: int = 1, detach: bool = True) -> Tensor:
    """da_4dvar.py:666-681.  x (69,nlat,nlon) physical -> physical."""
    z = ((x - c.mean) / c.std).unsqueeze(0)
    if c.lr is not None:
        z = F.interpolate(z, c.lr)                                       # :670-671
    for _ in range(steps):
        z = nets.flow(z)[:, :69]
        if detach:
            z = z.detach()
    if c.lr is not None:
        z = F.interpolate(z, tuple(x.shape[-2:]))                        # :678-679
    return z.reshape(x.shape) * c.std + c.mean


def _decode(z: Tensor, c: Case, nets) -> Tensor:
    """VAE_lr.decoder, or decoder_hr (nf_model/vae.py:87-90) when the analysis grid is finer than the network grid."""
    d = nets.decode(z)
    return d if c.lr is None else F.interpolate(d, tuple(c.xb.shape[-2:]))


def trajectory(z: Tensor, c: Case, nets) -> Tensor:
    """x_pred (T,69,nlat,nlon): x_0 = xb + D(z) stdTr sigma, x_{t+1} = M(x_t); da_4dvar.py:1185-1195."""
    x = (_decode(z, c, nets) * c.stdTr) * c.std.reshape(1, -1, 1, 1) + c.xb
    x = x[0]
    xs = [x]
    for _ in range(c.T - 1):
        x = integrate(x, c, nets, 1, detach=False)[:69]
        xs.append(x)
    return torch.stack(xs, 0)


def obs_interp_matrix(dim_in: int = 13, dim_out: int = 40) -> np.ndarray:
    """obs_interpolater.get_interp (da_4dvar.py:62-82): rows = `dim_out` pressure levels equally spaced in log p between 50 and
    1000 hPa (rounded), columns = the 13 model levels; linear interpolation in log p, float64 weights stored as float32."""
    level = [50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000]
    new = np.round(np.exp(np.linspace(3.91202301, 6.90775528, dim_out)))
    w = torch.zeros(dim_out, dim_in)
    for i in range(len(new)):
        for j in range(len(level)):
            if new[i] == level[j]:
                w[i, j] = 1
            elif j + 1 < len(level) and level[j] < new[i] < level[j + 1]:
                d = np.log(level[j + 1]) - np.log(level[j])
                w[i, j] = (np.log(level[j + 1]) - np.log(new[i])) / d
                w[i, j + 1] = (np.log(new[i]) - np.log(level[j])) / d
    return w.numpy()


def augment_levels(x_pred: Tensor, interp: Tensor, nlev: int = 13) -> Tensor:
    """(T,69,H,W) -> (T,4+5*dim_out,H,W): the four surface channels, then each of the five upper-air variables interpolated from its
    13 model levels to the observation levels (da_4dvar.py:1196-1206)."""
    parts = [x_pred[:, :4]]
    for i in range(5):
        mat = x_pred[:, 4 + i * nlev:4 + (i + 1) * nlev]
        parts.append(F.linear(mat.transpose(1, 3), interp).transpose(1, 3))
    return torch.cat(parts, 1)


def loss_terms(z, c, nets):
    """(J_reg, J_obs) with J = J_reg + obs_coeff J_obs; da_4dvar.py:1184,1207-1208."""
    j_reg = torch.sum(z ** 2) / 2
    x_pred = trajectory(z, c, nets)
    if c.interp is not None:
        x_pred = augment_levels(x_pred, c.interp)
    j_obs = torch.sum(c.H * (x_pred - c.yo) ** 2 / c.R) / 2
    return j_reg, j_obs


def cost_and_grad(z_np: np.ndarray, c: Case, nets):
    """One closure() call: returns (J, J_reg, J_obs, grad_z) as python floats / numpy."""
    z = torch.from_numpy(np.ascontiguousarray(z_np)).clone().to(c.xb.device).requires_grad_(True)
    j_reg, j_obs = loss_terms(z, c, nets)
    J = j_reg + c.obs_coeff * j_obs
    J.backward()
    return float(J.detach()), float(j_reg.detach()), float(j_obs.detach()), z.grad.detach().cpu().numpy()


# ---- metrics (utils/metrics.py) ------------------------------------------------------------
def _lat_weight(num_lat: int) -> Tensor:
    j = torch.arange(0, num_lat)
    lat = 90.0 - j * 180.0 / float(num_lat - 1)                          # metrics.py:5-6
    cosl = torch.cos(3.1416 / 180.0 * lat)                               # sic: 3.1416, metrics.py:10
    return (num_lat * cosl / torch.sum(cosl)).reshape(1, 1, -1, 1)


def wrmse(pred: Tensor, gt: Tensor, data_std: Tensor) -> Tensor:
    """Metrics.WRMSE: (n,c,h,w) normalised fields -> (c,) physical units; metrics.py:282-296,544."""
    w = _lat_weight(pred.shape[2])
    return torch.mean(torch.sqrt(torch.mean(w * (pred - gt) ** 2.0, dim=(-1, -2))), dim=0) * data_std


def bias(pred: Tensor, gt: Tensor, data_std: Tensor) -> Tensor:
    """Metrics.Bias: metrics.py:65-82,265-267,473-474."""
    w = _lat_weight(pred.shape[2])
    return torch.mean(torch.mean(w * (pred - gt), dim=(-1, -2)), dim=0) * data_std


def analysis(z: Tensor, c: Case, nets) -> Tensor:
    """xhat (69,nlat,nlon) physical; da_4dvar.py:1256-1259, 1301-1306."""
    with torch.no_grad():
        out = _decode(z, c, nets)
        return out[0] * c.stdTr[0] * c.std + c.xb


def diagnostics(z, c, nets):
    """(WRMSE[69], Bias[69]) of the current analysis against gt[0]; da_4dvar.py:1256-1264."""
    xhat = analysis(z, c, nets)
    xn = ((xhat - c.mean) / c.std).unsqueeze(0)
    gn = ((c.gt[0] - c.mean) / c.std).unsqueeze(0)
    return wrmse(xn, gn, c.std64), bias(xn, gn, c.std64)


def one_step_da(c: Case, nets, nit: int = 1, max_iter: int = 10,
                latent: int = 32, z0: Optional[np.ndarray] = None, log=None):
    """The vae4dvar branch of one_step_DA (da_4dvar.py:1238-1306) with torch.optim.LBFGS.

    Returns dict(xa, z, bg_wrmse, ana_wrmse, bg_bias, ana_bias, J_history, n_evals).
    """
    nlat, nlon = c.xb.shape[-2:] if c.lr is None else c.lr
    z = torch.zeros(1, latent, nlat, nlon) if z0 is None else torch.from_numpy(z0).clone()
    z.requires_grad_(True)
    opt = torch.optim.LBFGS([z], history_size=10, max_iter=max_iter, line_search_fn="strong_wolfe")
    evals = {"n": 0}
    hist = []

    def closure():
        opt.zero_grad()
        j_reg, j_obs = loss_terms(z, c, nets)
        J = j_reg + c.obs_coeff * j_obs
        J.backward()
        evals["n"] += 1
        hist.append(float(J))
        return J

    out = {}
    for kk in range(nit + 1):
        w, b = diagnostics(z.detach(), c, nets)
        if kk == 0:
            out["bg_wrmse"], out["bg_bias"] = w.numpy(), b.numpy()
        if kk == nit:
            out["ana_wrmse"], out["ana_bias"] = w.numpy(), b.numpy()
        if log:
            log(kk, w, b)
        if kk < nit:
            opt.step(closure)
    out["xa"] = analysis(z.detach(), c, nets).numpy()
    out["z"] = z.detach().numpy()
    out["J_history"] = np.asarray(hist)
    out["n_evals"] = evals["n"]
    return out
