"""BASELINE.json configs[4]: cycled data assimilation -- `--cycles` consecutive analysis-forecast cycles per
chain (da_4dvar.py:1314-1342: observations of the window, one_step_DA with Nit = 4 x LBFGS.step(max_iter=10), diagnostics, save,
forecast to the next window start), one independent chain per GPU (different truth / background / mask seeds), no data-path
collective; at the end one NCCL sum of the metric accumulator and of the per-rank timings.
    python tools/run_cycles.py --cycles 30
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 tools/run_cycles.py --cycles 30
At 128x256 the forecast operator of the cycle is the flow model (the 0.25-degree LGUnet_all_1 needs the 721x1440 grid); stated in
the output.  Rank 0 prints one JSON line (cycles/hour over the whole job, per-rank seconds, WRMSE of the first / last cycle)."""
import argparse
import json
import os
import pathlib
import sys
import tempfile

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import torch.distributed as dist

from vaevar_b200.config import DECODER_FULL, FLOW_FULL, era5_stats, small
from vaevar_b200.cycle import CycledDA, TwinObs
from vaevar_b200.da import VaeVar4D
from vaevar_b200.dist import MetricAccumulator
from vaevar_b200.synth import make_case, make_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--cycles", type=int, default=30)
ap.add_argument("--T", type=int, default=6)
ap.add_argument("--nit", type=int, default=4)
ap.add_argument("--obs-frac", type=float, default=0.10)
ap.add_argument("--small", action="store_true")
ap.add_argument("--out", default="")
a = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
dcfg, fcfg = (small(DECODER_FULL), small(FLOW_FULL)) if a.small else (DECODER_FULL, FLOW_FULL)
agent = VaeVar4D(dcfg, fcfg, make_state_dict(dcfg, seed=0), make_state_dict(fcfg, seed=1), da_win=a.T, Nit=a.nit, device=dev, verbose=False)
case = make_case(a.T, *dcfg.img_size, obs_frac=a.obs_frac, seed=100 + rank)           # chain `rank`: its own truth and background
with tempfile.TemporaryDirectory() as tmp:
    warm = CycledDA(agent, TwinObs(agent, torch.from_numpy(case["gt"][0]), obs_frac=a.obs_frac, seed=rank), torch.from_numpy(case["xb"]),
                    name=f"warm{rank}", root=tmp, n_cycles=1, resume=False)
    warm.run_assimilation()                          # graph capture, lazy loads: not part of the measured chain
    for v in agent.metrics_list.values():
        v.clear()
    agent.history.clear()
    run = CycledDA(agent, TwinObs(agent, torch.from_numpy(case["gt"][0]), obs_frac=a.obs_frac, seed=rank), torch.from_numpy(case["xb"]),
                   name=f"chain{rank}", root=tmp, n_cycles=a.cycles, resume=False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    r = run.run_assimilation()
    torch.cuda.synchronize()
    files = sorted(p.name for p in (pathlib.Path(tmp) / f"chain{rank}").iterdir())
acc = MetricAccumulator(69, dev)
bg = torch.stack(agent.metrics_list["bg_wrmse"]); an = torch.stack(agent.metrics_list["ana_wrmse"]); bi = torch.stack(agent.metrics_list["ana_bias"])
for k in range(an.shape[0]):
    acc.add(float(agent.history[(k + 1) * a.nit - 1]["loss"]), float(agent.history[(k + 1) * a.nit - 1]["gmax"]), an[k], bi[k])
secs = torch.zeros(world, dtype=torch.float64, device=dev); secs[rank] = sum(run.cycle_seconds)
z500 = torch.zeros(world, 4, dtype=torch.float64, device=dev)
z500[rank] = torch.tensor([float(bg[0, 11]), float(an[0, 11]), float(bg[-1, 11]), float(an[-1, 11])], dtype=torch.float64)
evals = torch.zeros(world, dtype=torch.float64, device=dev); evals[rank] = sum(h["func_evals"] if "func_evals" in h else h["n_evals"] for h in agent.history[-a.nit:])
if world > 1:
    for t in (secs, z500, evals):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
acc.reduce()
if rank == 0:
    s = acc.summary()
    per = secs.tolist()
    out = {"config": "BASELINE.json configs[4]: cycled DA, %d consecutive analysis-forecast cycles per chain, %d chains (one per GPU)" % (a.cycles, world),
           "T": a.T, "nit": a.nit, "obs_frac": a.obs_frac, "small": a.small, "world": world, "cycles_per_chain": a.cycles,
           "cycles_total": s["n_cases"], "seconds_per_rank": per, "imbalance": max(per) / max(min(per), 1e-9),
           "seconds_per_cycle": max(per) / a.cycles, "da_cycles_per_hour": 3600.0 * s["n_cases"] / max(per),
           "da_cycles_per_hour_per_gpu": 3600.0 * a.cycles / max(per),
           "forecast_operator": "flow model on the 128x256 engine grid (da_4dvar.py:1329 uses LGUnet_all_1 at 721x1440)",
           "z500_wrmse_per_chain[bg first, ana first, bg last, ana last]": z500.tolist(), "rms_ana_wrmse_z500": s["rms_wrmse"][11],
           "mean_J_final": s["mean_J"], "files_per_chain": files, "func_evals_last_cycle_rank0": float(evals[0])}
    if a.out:
        pathlib.Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        pathlib.Path(a.out).write_text(json.dumps(out))
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
