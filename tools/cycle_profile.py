"""Where one analysis-forecast cycle spends its wall-clock time (host + device), phase by phase.
    [OMP_NUM_THREADS=1] python tools/cycle_profile.py [--T 6]"""
import argparse
import pathlib
import sys
import tempfile
import time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from bench import build_inputs
from vaevar_b200.cycle import CycledDA, TwinObs
from vaevar_b200.da import VaeVar4D
from vaevar_b200.engine import LBFGS

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=6)
a = ap.parse_args()
dcfg, fcfg, sd_d, sd_f, case = build_inputs(a.T, 0.10, 0)
agent = VaeVar4D(dcfg, fcfg, sd_d, sd_f, da_win=a.T, Nit=4, verbose=False)
eng = agent.engine
obs = TwinObs(agent, torch.from_numpy(case["gt"][0]), obs_frac=0.1, seed=0)
yo, H, R, gt = obs.window(0)
xb = torch.from_numpy(case["xb"]).cuda()
sync = torch.cuda.synchronize


def timed(name, fn):
    sync(); t0 = time.time(); r = fn(); sync()
    print(f"{name:28s} {1e3 * (time.time() - t0):9.2f} ms", flush=True)
    return r


for rep in range(2):
    print(f"--- pass {rep}")
    timed("set_case", lambda: eng.set_case(xb, yo, H, R, 1.0))
    z = torch.zeros(1, 32, 128, 256, device="cuda")
    opt = LBFGS(eng, 10, 10)
    timed("decode", lambda: eng.decode(z))
    xh = eng.decode(z)
    timed("metrics", lambda: eng.metrics(xh, gt[0]))
    timed("cost (+.cpu())", lambda: eng.cost(z).cpu())
    timed("cost_grad x1", lambda: eng.cost_grad(z))
    for k in range(4):
        info = timed(f"lbfgs.step {k}", lambda: opt.step(z))
        print("      evals", info["n_evals"], "-> per eval %.2f ms" % 0.0)
    timed("integrate(1)", lambda: eng.integrate(eng.decode(z), 1))
    with tempfile.TemporaryDirectory() as tmp:
        timed("np.save xb", lambda: np.save(pathlib.Path(tmp) / "xb", eng.decode(z).cpu().numpy()))
    opt.close()
    with tempfile.TemporaryDirectory() as tmp:
        run = CycledDA(agent, obs, xb, name="p", root=tmp, n_cycles=1, resume=False)
        timed("CycledDA 1 cycle (total)", run.run_assimilation)
