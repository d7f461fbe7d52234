"""Host-side logic on CPU: configs, synthetic generator conventions, case sharding and the gloo metric reduction."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vaevar_b200.config import DECODER_FULL, ENCODER_FULL, FLOW_FULL, era5_stats
from vaevar_b200.dist import MetricAccumulator, shard_cases
from vaevar_b200.synth import make_case, obs_variance


def test_channel_shuffle_of_the_output_halves():
    # transformer.py:616-623: all first halves, then all second halves
    m = DECODER_FULL.output_channel_map()
    assert len(m) == 69 and DECODER_FULL.mean_half == 32
    assert m[:3] == [(0, 0), (0, 1), (1, 0)] and m[32] == (0, 2) and m[34] == (1, 6)
    assert FLOW_FULL.mean_half == 69 and FLOW_FULL.out_chans == 138
    assert ENCODER_FULL.mean_half == 32 and ENCODER_FULL.out_chans == 64


def test_constants():
    mean, std, stdtr = era5_stats()
    assert mean.shape == std.shape == stdtr.shape == (69,)
    assert abs(mean[3] - 100980.83590625007) < 1e-6 and abs(std[0] - 5.610453475051704) < 1e-12
    assert abs(stdtr[17] - 0.50658824) < 1e-9


def test_obs_variance_modify_tp():
    _, std, _ = era5_stats()
    v = obs_variance(0.005, 2)
    base = (np.float32(0.005) ** 2) * std.astype(np.float32) ** 2
    np.testing.assert_allclose(v[:2], base[:2], rtol=1e-6)
    np.testing.assert_allclose(v[2], base[2] / 16, rtol=1e-6)
    np.testing.assert_allclose(v[56:], base[56:] / 16, rtol=1e-6)


def test_case_conventions():
    c = make_case(3, 32, 64, obs_frac=0.1, seed=1)
    H = c["H"]
    assert H.shape == (3, 69, 32, 64) and set(np.unique(H)) == {0.0, 1.0}
    assert (H == H[0, 0]).all()                      # same columns for all channels and all times (da_4dvar.py:282-292)
    assert int(H[0, 0].sum()) == int(0.1 * 32 * 64)
    assert np.array_equal(c["yo"], c["gt"])           # noise-free obs (da_4dvar.py:449)
    assert (c["R"][1] == c["R"][0]).all()             # Q = 0 (da_4dvar.py:540-541)


def test_shard_cases_partition():
    for world in (1, 2, 4, 8):
        parts = [shard_cases(64, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(64))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    acc = MetricAccumulator(4)
    for case in shard_cases(5, rank, world):
        acc.add(float(case), 1.0, torch.full((4,), float(case)), torch.full((4,), 1.0))
    acc.reduce()
    q.put((rank, acc.buf.clone()))
    dist.destroy_process_group()


def test_metric_reduction_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29531, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    serial = MetricAccumulator(4)
    for case in range(5):
        serial.add(float(case), 1.0, torch.full((4,), float(case)), torch.full((4,), 1.0))
    for r in range(2):
        assert torch.equal(got[r], serial.buf)
