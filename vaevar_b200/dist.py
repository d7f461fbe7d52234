"""Multi-GPU = independent assimilation cases: one process per GPU, each with its own engine
and a full replica of both networks; no collective on the data path.  The only exchange is a sum-reduction of a small
float64 metric accumulator (count, sum J, sum |grad J|, per-channel sum WRMSE^2 / bias), the pattern of
utils/misc.py:33-45 in the reference's training code."""
from __future__ import annotations

from typing import List

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
