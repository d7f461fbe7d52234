"""ms per cost+gradient on the reference's REAL geometry: analysis grid 69 x 721 x 1440 over the 128 x 256 network grid
(decoder_hr + integrate(interpolation=True); vv_set_case_native), next to the network-grid number of the same engine.
    python tools/native_bench.py [--T 6] [--reps 20] > profiles/r1_native_T6.json"""
import argparse
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from vaevar_b200.config import DECODER_FULL, FLOW_FULL, era5_stats
from vaevar_b200.engine import Engine
from vaevar_b200.synth import make_state_dict, obs_variance

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=6)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--frac", type=float, default=0.10)
a = ap.parse_args()
dev = "cuda:0"
T, hr, lr = a.T, (721, 1440), (128, 256)
e = Engine(DECODER_FULL, FLOW_FULL, T=T)
e.load_state_dict(0, make_state_dict(DECODER_FULL, seed=0)); e.load_state_dict(1, make_state_dict(FLOW_FULL, seed=1)); e.finalize()
mean, std, _ = era5_stats()
m = torch.from_numpy(mean).float().to(dev).reshape(1, 69, 1, 1); s = torch.from_numpy(std).float().to(dev).reshape(1, 69, 1, 1)
gen = torch.Generator(device=dev).manual_seed(0)


def case(grid):
    gt = m + s * torch.randn(T, 69, *grid, device=dev, generator=gen)
    xb = gt[0] + 0.1 * s[0] * torch.randn(69, *grid, device=dev, generator=gen)
    mask = torch.zeros(grid[0] * grid[1], device=dev)
    mask[torch.randperm(grid[0] * grid[1], device=dev, generator=gen)[: int(a.frac * grid[0] * grid[1])]] = 1.0
    H = mask.reshape(1, 1, *grid).expand(T, 69, *grid).contiguous()
    R = torch.from_numpy(obs_variance(0.005, 2)).float().to(dev).reshape(1, 69, 1, 1).expand(T, 69, *grid).contiguous()
    return xb, gt, H, R


z = 0.1 * torch.randn(1, 32, *lr, device=dev, generator=gen)
J = torch.empty(3, dtype=torch.float64, device=dev); g = torch.empty_like(z)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit():
    for _ in range(4):
        e.cost_grad(z, J, g)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0.record(); e.cost_grad(z, J, g); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


out = {"T": T, "obs_frac": a.frac, "reps": a.reps}
xb, gt, H, R = case(lr)
e.set_case(xb, gt, H, R, 1.0)
out["network_grid"] = {"grid": lr, "n_obs": e.n_obs, "ms_per_cost_grad": timeit(), "launches": e.last_launch_count, "J": float(J[0])}
del xb, gt, H, R
xb, gt, H, R = case(hr)
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); e.set_case_native(xb, gt, H, R, 1.0); t1.record(); torch.cuda.synchronize()
out["native"] = {"grid": hr, "n_obs": e.n_obs, "set_case_ms": t0.elapsed_time(t1), "ms_per_cost_grad": timeit(), "launches": e.last_launch_count,
                 "J": float(J[0]), "grad_finite": bool(torch.isfinite(g).all()), "hbm_used_gb": torch.cuda.max_memory_allocated() / 1e9}
out["native_over_network_grid"] = out["native"]["ms_per_cost_grad"] / out["network_grid"]["ms_per_cost_grad"]
print(json.dumps(out))
