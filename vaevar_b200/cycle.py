"""Cycled data assimilation around the engine: the driver of da_4dvar.py:1314-1342 restricted to da_mode == "vae4dvar".

    cyclic_4dvar.run_assimilation    da_4dvar.py:1314-1342
    get_current_states / save_ckpt   da_4dvar.py:683-702   (resume files xb.npy + current_time.txt)
    save_eval_result / load_eval_ckpts  da_4dvar.py:704-727   (metric .npy dumps)
    get_obs_info ("free" observations)  da_4dvar.py:758-805, 276-292, 442-450

The reference fetches the truth from an S3 store (data_reader.get_state, da_4dvar.py:148-166), which does not exist offline, so the
observation source is pluggable (`ObsSource`).  The forecast step is `agent.forecast_model` (an `LGUnet_all_1` shell on the analysis
grid, da_4dvar.py:484, 1329) when the agent has one, else the flow model applied on the engine grid.  `TwinObs` is the identical-twin source the synthetic
configs use: a truth run advanced by the same flow model, observed noise-free through a fixed random column mask.
"""
from __future__ import annotations

import os
import pathlib
import time
from typing import Dict, List, Optional, Protocol, Tuple

import numpy as np
import torch

from .da import VaeVar4D
from .synth import obs_variance


class ObsSource(Protocol):
    def window(self, cycle: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """(yo, H, R, gt), each (T,69,nlat,nlon) float32 on the agent's device, for the window that starts at `cycle`."""


class TwinObs:
    """Identical-twin observations: truth_{k+1} = M^{steps}(truth_k); yo = gt (noise-free, da_4dvar.py:449) through a random
    column mask shared by all channels and all T (da_4dvar.py:282-292); R = (obs_std sigma_c)^2 with modify_tp (:106-127)."""

    def __init__(self, agent: VaeVar4D, truth0: torch.Tensor, obs_frac: float = 0.10, seed: int = 0, obs_std: float = 0.005,
                 modify_tp: int = 2, steps_per_cycle: int = 1):
        self.agent, self.steps = agent, steps_per_cycle
        self.truth = truth0.to(agent.device, torch.float32)
        self._cycle = 0
        # the observation grid is the truth's: the network grid, or a finer analysis grid (the reference's 721x1440)
        T, C, nlat, nlon = agent.da_win, agent.nchannel, int(truth0.shape[-2]), int(truth0.shape[-1])
        self.native = (nlat, nlon) != (agent.nlat, agent.nlon)
        rng = np.random.Generator(np.random.PCG64(2000 + seed))
        cols = rng.choice(nlat * nlon, int(obs_frac * nlat * nlon), replace=False)
        mask = np.zeros(nlat * nlon, np.float32)
        mask[cols] = 1.0
        self.H = torch.from_numpy(np.broadcast_to(mask.reshape(1, 1, nlat, nlon), (T, C, nlat, nlon)).copy()).to(agent.device)
        var = obs_variance(obs_std, modify_tp).reshape(1, C, 1, 1)
        self.R = torch.from_numpy(np.broadcast_to(var, (T, C, nlat, nlon)).copy()).to(agent.device)

    def truth_window(self, cycle: int) -> torch.Tensor:
        native = getattr(self, "native", False)
        while self._cycle < cycle:                         # advance the truth run to the start of this window: the cycle's own forecast
            self.truth = self.agent.integrate(self.truth, getattr(self.agent, "forecast_model", None), self.steps)      # operator (identical twin)
            self._cycle += 1
        gt = [self.truth]
        for _ in range(self.agent.da_win - 1):             # inside the window: the flow model, as the closure applies it (da_4dvar.py:1191)
            gt.append(self.agent.integrate(gt[-1], None, 1, interpolation=native))
        return torch.stack(gt)

    def window(self, cycle: int):
        gt = self.truth_window(cycle)
        return gt.clone(), self.H, self.R, gt


def augment_levels(x: torch.Tensor, interp: torch.Tensor, nlev: int = 13) -> torch.Tensor:
    """(T,69,H,W) -> (T,4+5*dim_out,H,W): surface channels, then every upper-air variable interpolated from its model levels to the
    observation levels (da_4dvar.py:770-776, 747-754, 1196-1206).  Data preparation, not the hot path: plain torch on x's device."""
    parts = [x[:, :4]]
    for i in range(5):
        parts.append(torch.einsum("ol,tlhw->tohw", interp.to(x), x[:, 4 + i * nlev:4 + (i + 1) * nlev]))
    return torch.cat(parts, 1)


class RealSimuObs(TwinObs):
    """The "real_simu" observation branch of get_obs_info (da_4dvar.py:766-800) over the identical-twin truth: observations live in
    the augmented space of `agent.obs_interp` (4 + 5 x 40 channels), each augmented channel and time level has its own sparse
    mask (the offline stand-in for data_reader.get_real_obs' station / sounding locations), an optional quality-control filter keeps
    |yo_real - aug(gt)| < filter_coeff * std_layer_aug (:780-787), yo = aug(gt) * H (:796-797) and R = aug(R_static)
    (get_R_matrix_from_gt, :745-756).  Use with VaeVar4D(obs_type="real_simu")."""

    def __init__(self, agent: VaeVar4D, truth0: torch.Tensor, obs_frac: float = 0.02, seed: int = 0, obs_std: float = 0.005,
                 modify_tp: int = 2, steps_per_cycle: int = 1, filter_coeff: Optional[float] = None, std_layer_aug=None, yo_real=None):
        self.agent, self.steps = agent, steps_per_cycle
        self.truth = truth0.to(agent.device, torch.float32)
        self._cycle = 0
        self.obs_frac, self.seed = obs_frac, seed
        self.filter_coeff, self.std_layer_aug, self.yo_real = filter_coeff, std_layer_aug, yo_real
        var = torch.from_numpy(obs_variance(obs_std, modify_tp).astype(np.float32)).reshape(1, agent.nchannel, 1, 1).to(agent.device)
        self.R_aug = augment_levels(var.expand(agent.da_win, -1, -1, -1), agent.obs_interp.interp)       # (T, A, 1, 1)

    def window(self, cycle: int):
        gt = self.truth_window(cycle)
        gt_aug = augment_levels(gt, self.agent.obs_interp.interp)
        gen = torch.Generator(device=gt.device).manual_seed(3000 + 977 * self.seed + cycle)
        H = (torch.rand(gt_aug.shape, device=gt.device, generator=gen) < self.obs_frac).float()
        if self.filter_coeff is not None and self.yo_real is not None:
            lim = self.filter_coeff * torch.as_tensor(self.std_layer_aug, dtype=torch.float32, device=gt.device).reshape(1, -1, 1, 1)
            d = self.yo_real(cycle).to(gt_aug) - gt_aug
            H = H * ((d < lim) & (d > -lim)).float()
        return gt_aug * H, H, self.R_aug.expand_as(gt_aug).contiguous(), gt


class CycledDA:
    """run_assimilation() with resume: every cycle = get_obs_info -> one_step_DA -> save -> integrate (forecast)."""

    def __init__(self, agent: VaeVar4D, obs: ObsSource, xb0: torch.Tensor, name: str = "vaevar_b200", root: str = "da_cycle_results",
                 n_cycles: int = 30, save_interval: int = 1, forecast_steps: int = 1, resume: bool = True, save_field: bool = False):
        self.agent, self.obs, self.name = agent, obs, name
        self.dir = pathlib.Path(root) / name
        self.dir.mkdir(parents=True, exist_ok=True)
        self.n_cycles, self.save_interval, self.forecast_steps, self.save_field = n_cycles, save_interval, forecast_steps, save_field
        self.cycle_seconds: List[float] = []
        self.current_cycle, self.xb = self.get_current_states(xb0) if resume else (0, xb0)
        self.xb = self.xb.to(agent.device, torch.float32)
        if resume:
            self.load_eval_ckpts()

    # da_4dvar.py:683-696
    def get_current_states(self, xb0: torch.Tensor):
        f = self.dir / "current_time.txt"
        cycle = int(f.read_text()) if f.exists() else 0
        x = self.dir / "xb.npy"
        xb = torch.from_numpy(np.load(x)) if x.exists() else xb0
        return cycle, xb

    # da_4dvar.py:698-702
    def save_ckpt(self):
        np.save(self.dir / "xb", self.xb.cpu().numpy())
        (self.dir / "current_time.txt").write_text(str(self.current_cycle))

    # da_4dvar.py:704-722
    def save_eval_result(self, xa: Optional[torch.Tensor] = None):
        for key, vals in self.agent.metrics_list.items():
            np.save(self.dir / key, np.stack([np.asarray(v, np.float64) for v in vals]) if vals else np.zeros((0,)))
        if self.save_field and xa is not None:
            np.save(self.dir / f"xb_{self.current_cycle}", self.xb.cpu().numpy())
            np.save(self.dir / f"xa_{self.current_cycle}", xa.cpu().numpy())

    # da_4dvar.py:724-727
    def load_eval_ckpts(self):
        for key in self.agent.metrics_list:
            f = self.dir / f"{key}.npy"
            if f.exists():
                self.agent.metrics_list[key] = [torch.from_numpy(r) for r in np.load(f)]

    # da_4dvar.py:1314-1342
    def run_assimilation(self) -> Dict[str, float]:
        a = self.agent
        epoch = 0
        while self.current_cycle < self.n_cycles:
            yo, H, R, gt = self.obs.window(self.current_cycle)          # the data source is not part of the timed cycle
            if torch.device(a.device).type == "cuda":
                torch.cuda.current_stream().synchronize()     # this chain only: other chains may share the GPU
            t0 = time.time()
            xa = a.one_step_DA(gt, self.xb, yo, H, R, "vae4dvar")
            self.save_eval_result(xa)
            self.xb = a.integrate(xa, getattr(a, "forecast_model", None), self.forecast_steps)     # da_4dvar.py:1329
            if torch.device(a.device).type == "cuda":
                torch.cuda.current_stream().synchronize()     # this chain only: other chains may share the GPU
            self.current_cycle += 1
            if epoch % self.save_interval == 0:
                self.save_ckpt()
            epoch += 1
            self.cycle_seconds.append(time.time() - t0)
        self.save_eval_result()
        n = max(len(self.cycle_seconds), 1)
        return {"cycles": len(self.cycle_seconds), "seconds_per_cycle": sum(self.cycle_seconds) / n,
                "cycles_per_hour": 3600.0 * n / max(sum(self.cycle_seconds), 1e-9)}
