"""Multi-GPU = independent assimilation cases: one process per GPU, each with its own engine
and a full replica of both networks; no collective on the data path.  The only exchange is a sum-reduction of a small
float64 metric accumulator (count, sum J, sum |grad J|, per-channel sum WRMSE^2 / bias), the pattern of
utils/misc.py:33-45 in the reference's training code."""
from __future__ import annotations

import contextlib
import threading
import time
from typing import Callable, Dict, List

import torch
import torch.distributed as dist


def shard_cases(n_cases: int, rank: int, world: int) -> List[int]:
    """case i -> rank i mod world (round-robin keeps the per-rank counts within one of each other)."""
    return list(range(rank, n_cases, world))


class MetricAccumulator:
    """float64[3 + 2*C] = [n_cases, sum J_final, sum |grad|_inf, sum_c WRMSE_c^2 ..., sum_c bias_c ...]."""

    def __init__(self, n_channels: int = 69, device="cpu"):
        self.C = n_channels
        self.buf = torch.zeros(3 + 2 * n_channels, dtype=torch.float64, device=device)

    def add(self, J_final: float, gmax: float, wrmse: torch.Tensor, bias: torch.Tensor):
        self.buf[0] += 1
        self.buf[1] += J_final
        self.buf[2] += gmax
        self.buf[3:3 + self.C] += wrmse.to(self.buf).double() ** 2
        self.buf[3 + self.C:] += bias.to(self.buf).double()

    def reduce(self) -> "MetricAccumulator":
        """Sum over ranks (NCCL on GPUs, gloo in the CPU tests); a no-op without an initialised process group."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
        return self

    def summary(self):
        n = max(float(self.buf[0]), 1.0)
        return {"n_cases": int(self.buf[0]), "mean_J": float(self.buf[1]) / n, "mean_gmax": float(self.buf[2]) / n,
                "rms_wrmse": (self.buf[3:3 + self.C] / n).sqrt().tolist(), "mean_bias": (self.buf[3 + self.C:] / n).tolist()}


def run_cases(agent, n_cases: int, make_case_fn: Callable[[int], Dict], rank: int = 0, world: int = 1, device="cpu") -> Dict:
    """`n_cases` independent assimilation cases (seeds 0..n-1), case i on rank i mod world, each one
    `agent.one_step_DA(gt, xb, yo, H, R)`; per-case results go into a MetricAccumulator that is summed over ranks once at the end;
    the elapsed time is the maximum over ranks.  Returns the summary with `cases_per_hour` (identical on every rank).

    `agent` may be a LIST of agents on the same device ("one or more cases per GPU"): this rank's cases are then dealt
    round-robin to one host thread per agent, each with its own engine, CUDA stream and launch graph, so the kernels of independent
    cases interleave on the GPU (a cost evaluation is a chain of ~1600 short kernels whose boundaries and tails leave SMs idle).
    Every case is still computed by exactly one engine in the same order of operations: the per-case records do not change."""
    agents = list(agent) if isinstance(agent, (list, tuple)) else [agent]
    agent = agents[0]
    acc = MetricAccumulator(agent.nchannel, device)
    cuda = torch.device(device).type == "cuda"
    multi = dist.is_available() and dist.is_initialized() and world > 1
    if multi:
        dist.barrier()
    if cuda:
        torch.cuda.synchronize()
    t0 = time.time()
    mine = shard_cases(n_cases, rank, world)
    records = torch.zeros(n_cases, 3, dtype=torch.float64, device=device)       # per case: final J, z500 analysis WRMSE, checksum of xa
    results: Dict[int, tuple] = {}

    def work(ag, cases):
        if cuda:                                   # a new host thread starts on device 0 with the legacy default stream
            torch.cuda.set_device(torch.device(device))
        ctx = torch.cuda.stream(torch.cuda.Stream(device=device)) if (cuda and len(agents) > 1) else contextlib.nullcontext()
        with ctx:
            for i in cases:
                c = make_case_fn(i)
                xa = ag.one_step_DA(c["gt"], c["xb"], c["yo"], c["H"], c["R"], "vae4dvar")
                info = ag.history[-1]
                results[i] = (float(info["loss"]), float(info["gmax"]), ag.metrics_list["ana_wrmse"][-1], ag.metrics_list["ana_bias"][-1],
                              float(torch.as_tensor(xa).double().sum()))
            if cuda:
                torch.cuda.current_stream().synchronize()

    if len(agents) == 1:
        work(agent, mine)
    else:
        threads = [threading.Thread(target=work, args=(ag, mine[k::len(agents)])) for k, ag in enumerate(agents)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        missing = [i for i in mine if i not in results]
        if missing:
            raise RuntimeError(f"cases {missing} did not finish (a worker thread failed)")
    for i in mine:                                 # accumulate in case order: the sums do not depend on the thread interleaving
        loss, gmax, w, b, chk = results[i]
        acc.add(loss, gmax, w, b)
        records[i, 0] = loss
        records[i, 1] = float(w[min(11, agent.nchannel - 1)])      # z500 (da_4dvar.py:1253)
        records[i, 2] = chk
    if cuda:
        torch.cuda.synchronize()
    mine_s = time.time() - t0
    el = torch.tensor([mine_s], dtype=torch.float64, device=device)
    per_rank = torch.zeros(world, dtype=torch.float64, device=device)
    per_rank[rank] = mine_s
    if multi:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
        dist.all_reduce(records, op=dist.ReduceOp.SUM)          # every case was written by exactly one rank
    acc.reduce()
    out = acc.summary()
    pr = per_rank.tolist()
    out.update(cases_in_flight_per_gpu=len(agents), seconds=float(el), cases_per_hour=3600.0 * out["n_cases"] / max(float(el), 1e-9), world=world, cases_on_this_rank=len(mine),
               seconds_per_rank=pr, imbalance=(max(pr) / max(min(pr), 1e-9)) if pr else 1.0,
               case_records=[[float(v) for v in r] for r in records.cpu()])
    return out
