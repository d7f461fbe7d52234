"""N independent assimilation cases (seeds 0..N-1) sharded over the GPUs of one node, one process per GPU,
no data-path collective; one NCCL sum of the metric accumulator and one max of the elapsed time at the end.
    python tools/run_cases.py --cases 8 [--T 6] [--nit 1] [--small]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/run_cases.py --cases 64
Rank 0 prints one JSON line (cases/hour over the whole job)."""
import argparse
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist

from vaevar_b200.config import DECODER_FULL, FLOW_FULL, small
from vaevar_b200.da import VaeVar4D
from vaevar_b200.dist import run_cases
from vaevar_b200.synth import make_case, make_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=8)
ap.add_argument("--T", type=int, default=6)
ap.add_argument("--nit", type=int, default=1)
ap.add_argument("--obs-frac", type=float, default=0.10)
ap.add_argument("--small", action="store_true", help="shrunken networks on a 32x64 grid (smoke test)")
ap.add_argument("--in-flight", type=int, default=1, help="independent cases in flight per GPU (one engine, stream and host thread each)")
ap.add_argument("--out", default="", help="write the JSON result (with the per-case records) to this file as well")
ap.add_argument("--check", default="", help="a result file of another run (e.g. 1 GPU): the per-case records must be bit-identical")
a = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
dcfg, fcfg = (small(DECODER_FULL), small(FLOW_FULL)) if a.small else (DECODER_FULL, FLOW_FULL)
sd_d, sd_f = make_state_dict(dcfg, seed=0), make_state_dict(fcfg, seed=1) if a.T > 1 else None
agents = [VaeVar4D(dcfg, fcfg if a.T > 1 else None, sd_d, sd_f, da_win=a.T, Nit=a.nit, device=dev, verbose=False) for _ in range(a.in_flight)]
mk = lambda i: make_case(a.T, *dcfg.img_size, obs_frac=a.obs_frac, seed=i)
c0 = mk(0)
for agent in agents:
    agent.one_step_DA(c0["gt"], c0["xb"], c0["yo"], c0["H"], c0["R"], "vae4dvar")      # warm-up: graph capture, lazy loads
    for v in agent.metrics_list.values():
        v.clear()
    agent.history.clear()
r = run_cases(agents if len(agents) > 1 else agents[0], a.cases, mk, rank, world, dev)
if rank == 0:
    r.update(T=a.T, nit=a.nit, obs_frac=a.obs_frac, small=a.small, rms_wrmse_z500=r["rms_wrmse"][11], config="BASELINE.json configs[3]")
    r.pop("rms_wrmse"); r.pop("mean_bias")
    if a.check:          # "N cases on N GPUs == the same cases on 1 GPU", per case, bit for bit
        other = json.loads(pathlib.Path(a.check).read_text())["case_records"]
        n = min(len(other), len(r["case_records"]))
        r["identical_to"] = {"file": a.check, "cases_compared": n, "bit_identical": r["case_records"][:n] == other[:n]}
    if a.out:
        pathlib.Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        pathlib.Path(a.out).write_text(json.dumps(r))
    brief = dict(r); brief["case_records"] = brief["case_records"][:2] + ["..."]
    print(json.dumps(brief))
if world > 1:
    dist.destroy_process_group()
