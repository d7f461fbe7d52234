"""BASELINE.json configs[4]: cycled data assimilation -- `--cycles` consecutive analysis-forecast cycles per
chain (da_4dvar.py:1314-1342: observations of the window, one_step_DA with Nit = 4 x LBFGS.step(max_iter=10), diagnostics, save,
forecast to the next window start), one or more independent chains per GPU (different truth / background / mask seeds; with
--chains-per-gpu S every chain has its own engine, CUDA stream and host thread, so their kernels interleave), no data-path
collective; at the end one NCCL sum of the metric accumulator and of the per-rank timings.
    python tools/run_cycles.py --cycles 30
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 tools/run_cycles.py --cycles 30
At 128x256 the forecast operator of the cycle is the flow model (the 0.25-degree LGUnet_all_1 needs the 721x1440 grid); stated in
the output.  Rank 0 prints one JSON line (cycles/hour over the whole job, per-rank seconds, WRMSE of the first / last cycle)."""
import argparse
import contextlib
import json
import os
import pathlib
import sys
import tempfile
import threading
import time

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
ap.add_argument("--chains-per-gpu", type=int, default=1, help="independent cycle chains in flight on every GPU (one engine, stream and host thread each)")
ap.add_argument("--native", action="store_true", help="the reference's real geometry: fields, observations and analysis on the 69x721x1440 grid "
                "(decoder_hr / integrate(interpolation=True) composed into the engine) and the forecast step by LGUnet_all_1 at 721x1440")
ap.add_argument("--chain-base", type=int, default=0, help="first chain id of this job (chain ids seed the truth, background and mask)")
ap.add_argument("--verbose", action="store_true", help="print the per-cycle z500 WRMSE of every chain of rank 0")
ap.add_argument("--out", default="")
a = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
dcfg, fcfg = (small(DECODER_FULL), small(FLOW_FULL)) if a.small else (DECODER_FULL, FLOW_FULL)
S = a.chains_per_gpu
sd_d, sd_f = make_state_dict(dcfg, seed=0), make_state_dict(fcfg, seed=1)
agents = [VaeVar4D(dcfg, fcfg, sd_d, sd_f, da_win=a.T, Nit=a.nit, device=dev, verbose=False) for _ in range(S)]
grid = dcfg.img_size
if a.native:
    from vaevar_b200.config import FORECAST_FULL
    from vaevar_b200.modules import LGUnet_all_1
    grid = FORECAST_FULL.img_size
    fmodel = LGUnet_all_1(**FORECAST_FULL.to_reference_kwargs(), keep_out=69).to(dev).eval()        # random-init weights of the shipped architecture
    for ag in agents:
        ag.forecast_model = fmodel
tmpdir = tempfile.TemporaryDirectory()
tmp = tmpdir.name
runs, files = [None] * S, [None] * S
gate = threading.Barrier(S + 1)
errors = []


def chain_worker(k):
    """Chain `rank * S + k`: its own truth, background and mask seeds; warm-up cycle (graph capture, lazy loads), then the measured chain."""
    try:
        torch.cuda.set_device(local)
        cid = a.chain_base + rank * S + k
        agent = agents[k]
        with torch.cuda.stream(torch.cuda.Stream(device=dev)) if S > 1 else contextlib.nullcontext():
            case = make_case(1, *grid, obs_frac=a.obs_frac, seed=100 + cid) if a.native else make_case(a.T, *dcfg.img_size, obs_frac=a.obs_frac, seed=100 + cid)
            mk = lambda name, n: CycledDA(agent, TwinObs(agent, torch.from_numpy(case["gt"][0]), obs_frac=a.obs_frac, seed=cid),
                                          torch.from_numpy(case["xb"]), name=name, root=tmp, n_cycles=n, resume=False)
            mk(f"warm{cid}", 1).run_assimilation()
            for v in agent.metrics_list.values():
                v.clear()
            agent.history.clear()
            runs[k] = mk(f"chain{cid}", a.cycles)
            torch.cuda.current_stream().synchronize()
            gate.wait()
            runs[k].run_assimilation()
            torch.cuda.current_stream().synchronize()
            gate.wait()
            files[k] = sorted(p.name for p in (pathlib.Path(tmp) / f"chain{cid}").iterdir())
    except Exception as ex:
        errors.append(repr(ex))
        gate.abort()


threads = [threading.Thread(target=chain_worker, args=(k,)) for k in range(S)]
for t in threads:
    t.start()
gate.wait()                                          # every chain of this rank is warmed up
if world > 1:
    dist.barrier()
t0 = time.time()
gate.wait()
rank_seconds = time.time() - t0
for t in threads:
    t.join()
if errors:
    raise RuntimeError(errors[0])
acc = MetricAccumulator(69, dev)
secs = torch.zeros(world, dtype=torch.float64, device=dev); secs[rank] = rank_seconds
z500 = torch.zeros(world * S, 4, dtype=torch.float64, device=dev)
evals = torch.zeros(world * S, dtype=torch.float64, device=dev)
for k, agent in enumerate(agents):
    bg = torch.stack(agent.metrics_list["bg_wrmse"]); an = torch.stack(agent.metrics_list["ana_wrmse"]); bi = torch.stack(agent.metrics_list["ana_bias"])
    for c in range(an.shape[0]):
        h = agent.history[(c + 1) * a.nit - 1]
        acc.add(float(h["loss"]), float(h["gmax"]), an[c], bi[c])
    if a.verbose and rank == 0:
        print(f"chain {a.chain_base + k}: z500 WRMSE per cycle (bg -> ana): " + " ".join(f"{float(bg[c, 11]):.4g}->{float(an[c, 11]):.4g}" for c in range(an.shape[0])),
              flush=True)
        print(f"chain {a.chain_base + k}: closure evals per L-BFGS step: " + " ".join(str(h["n_evals"]) for h in agent.history), flush=True)
        print(f"chain {a.chain_base + k}: final J per step: " + " ".join(f"{h['loss']:.4g}" for h in agent.history), flush=True)
    z500[rank * S + k] = torch.tensor([float(bg[0, 11]), float(an[0, 11]), float(bg[-1, 11]), float(an[-1, 11])], dtype=torch.float64)
    evals[rank * S + k] = sum(h["n_evals"] for h in agent.history[-a.nit:])
if world > 1:
    for t in (secs, z500, evals):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
acc.reduce()
if rank == 0:
    s = acc.summary()
    per = secs.tolist()
    out = {"config": "BASELINE.json configs[4]: cycled DA, %d consecutive analysis-forecast cycles per chain, %d chains (%d in flight per GPU)"
                     % (a.cycles, world * S, S),
           "T": a.T, "nit": a.nit, "obs_frac": a.obs_frac, "small": a.small, "world": world, "chains_per_gpu": S, "cycles_per_chain": a.cycles,
           "cycles_total": s["n_cases"], "seconds_per_rank": per, "imbalance": max(per) / max(min(per), 1e-9),
           "seconds_per_cycle_per_chain": max(per) / a.cycles, "da_cycles_per_hour": 3600.0 * s["n_cases"] / max(per),
           "da_cycles_per_hour_per_gpu": 3600.0 * a.cycles * S / max(per),
           "analysis_grid": list(grid),
           "forecast_operator": ("LGUnet_all_1 at 721x1440 (da_4dvar.py:1329), random-init weights of the shipped architecture" if a.native else
                                 "flow model on the 128x256 engine grid (da_4dvar.py:1329 uses LGUnet_all_1 at 721x1440: --native)"),
           "z500_wrmse_per_chain[bg first, ana first, bg last, ana last]": z500.tolist(), "rms_ana_wrmse_z500": s["rms_wrmse"][11],
           "mean_J_final": s["mean_J"], "files_per_chain": files[0], "closure_evals_last_cycle_per_chain": evals.tolist()}
    if a.out:
        pathlib.Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        pathlib.Path(a.out).write_text(json.dumps(out))
    print(json.dumps(out))
tmpdir.cleanup()
if world > 1:
    dist.destroy_process_group()
