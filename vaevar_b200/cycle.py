"""Cycled data assimilation around the engine: the driver of da_4dvar.py:1314-1342 restricted to da_mode == "vae4dvar".

    cyclic_4dvar.run_assimilation    da_4dvar.py:1314-1342
    get_current_states / save_ckpt   da_4dvar.py:683-702   (resume files xb.npy + current_time.txt)
    save_eval_result / load_eval_ckpts  da_4dvar.py:704-727   (metric .npy dumps)
    get_obs_info ("free" observations)  da_4dvar.py:758-805, 276-292, 442-450

The reference fetches the truth from an S3 store (data_reader.get_state, da_4dvar.py:148-166) and forecasts with the 0.25-degree
LGUnet_all_1; neither exists offline, so here the observation source is pluggable (`ObsSource`) and the forecast step is the flow
model applied on the engine grid (stated, not hidden).  `TwinObs` is the identical-twin source the synthetic
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
        T, C, nlat, nlon = agent.da_win, agent.nchannel, agent.nlat, agent.nlon
        rng = np.random.Generator(np.random.PCG64(2000 + seed))
        cols = rng.choice(nlat * nlon, int(obs_frac * nlat * nlon), replace=False)
        mask = np.zeros(nlat * nlon, np.float32)
        mask[cols] = 1.0
        self.H = torch.from_numpy(np.broadcast_to(mask.reshape(1, 1, nlat, nlon), (T, C, nlat, nlon)).copy()).to(agent.device)
        var = obs_variance(obs_std, modify_tp).reshape(1, C, 1, 1)
        self.R = torch.from_numpy(np.broadcast_to(var, (T, C, nlat, nlon)).copy()).to(agent.device)

    def window(self, cycle: int):
        while self._cycle < cycle:                         # advance the truth run to the start of this window
            self.truth = self.agent.integrate(self.truth, None, self.steps)
            self._cycle += 1
        gt = [self.truth]
        for _ in range(self.agent.da_win - 1):
            gt.append(self.agent.integrate(gt[-1], None, 1))
        gt = torch.stack(gt)
        return gt.clone(), self.H, self.R, gt


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
                torch.cuda.synchronize()
            t0 = time.time()
            xa = a.one_step_DA(gt, self.xb, yo, H, R, "vae4dvar")
            self.save_eval_result(xa)
            self.xb = a.integrate(xa, None, self.forecast_steps)
            if torch.device(a.device).type == "cuda":
                torch.cuda.synchronize()
            self.current_cycle += 1
            if epoch % self.save_interval == 0:
                self.save_ckpt()
            epoch += 1
            self.cycle_seconds.append(time.time() - t0)
        self.save_eval_result()
        n = max(len(self.cycle_seconds), 1)
        return {"cycles": len(self.cycle_seconds), "seconds_per_cycle": sum(self.cycle_seconds) / n,
                "cycles_per_hour": 3600.0 * n / max(sum(self.cycle_seconds), 1e-9)}
