"""Closure-by-closure trace of the device L-BFGS on the shrunken T-step case next to torch.optim.LBFGS on the CPU oracle.
    python tools/lbfgs_debug.py [--T 3] [--nit 1]"""
import argparse
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch

import gpu_check
from oracle import cost as oc
from oracle.lgunet import to_torch
from vaevar_b200.engine import LBFGS
from vaevar_b200.synth import make_case

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=3)
ap.add_argument("--nit", type=int, default=1)
a = ap.parse_args()
e, ds, fs, sd_d, sd_f = gpu_check._engine_small(a.T)
case = make_case(a.T, *ds.img_size, obs_frac=0.10, seed=0)
e.set_case(case["xb"], case["yo"], case["H"], case["R"], 1.0)
z = torch.zeros(1, 32, *ds.img_size, device="cuda")
opt = LBFGS(e, 10, 10)
for _ in range(a.nit):
    info = opt.step(z)
    print("engine step info", info)
h, t = opt.history(), opt.steps()
nets = oc.OracleNets(to_torch(sd_d), ds, to_torch(sd_f), fs)
r = oc.one_step_da(oc.Case(case), nets, nit=a.nit, max_iter=10)
ho = r["J_history"]
for i in range(max(len(h), len(ho))):
    le = f"t={t[i]:.6e} J={h[i]:.9e}" if i < len(h) else " " * 34
    lo = f"J={ho[i]:.9e}" if i < len(ho) else ""
    print(f"{i:3d} engine {le} | oracle {lo}")
print("z finite:", bool(torch.isfinite(z).all()), "|z|max", float(z.abs().max()))
