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
    q.put((rank, acc.buf.tolist()))       # plain list: a tensor would travel as a shared-memory handle that dies with the worker
    dist.destroy_process_group()


def _cases_worker(rank, world, port, q):
    from vaevar_b200.dist import run_cases
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Agent(_FakeAgent):
        def one_step_DA(self, gt, xb, yo, H, R, mode="vae4dvar"):
            seed = float(xb.flatten()[0])
            self.metrics_list["ana_wrmse"].append(torch.full((self.nchannel,), seed))
            self.metrics_list["ana_bias"].append(torch.full((self.nchannel,), -seed))
            self.history.append({"loss": 10.0 + seed, "gmax": 0.5 * seed})
            return xb

    a = Agent()
    a.history = []
    mk = lambda i: {k: torch.full((1, 3, 4, 8), float(i)) for k in ("gt", "yo", "H", "R")} | {"xb": torch.full((3, 4, 8), float(i))}
    r = run_cases(a, 5, mk, rank, world, "cpu")
    q.put((rank, {k: r[k] for k in ("n_cases", "mean_J", "mean_gmax", "rms_wrmse", "mean_bias", "cases_on_this_rank", "world", "case_records",
                                    "seconds_per_rank", "imbalance")}))
    dist.destroy_process_group()


def test_run_cases_world2_gloo():
    """SURVEY 8(d) config 4 host logic: 5 cases over 2 ranks, results identical on both ranks and equal to the serial sums."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cases_worker, args=(r, 2, 29537, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert got[0]["cases_on_this_rank"] == 3 and got[1]["cases_on_this_rank"] == 2
    for r in range(2):
        g = got[r]
        assert g["n_cases"] == 5 and g["world"] == 2
        assert abs(g["mean_J"] - (10.0 + 2.0)) < 1e-12 and abs(g["mean_gmax"] - 1.0) < 1e-12
        assert np.allclose(g["rms_wrmse"], np.sqrt((0 + 1 + 4 + 9 + 16) / 5.0)) and np.allclose(g["mean_bias"], -2.0)
    assert got[0]["rms_wrmse"] == got[1]["rms_wrmse"]
    # per-case records (final J, z500 WRMSE, checksum of the analysis) are the same on every rank and are those of a serial run:
    # what `tools/run_cases.py --check` compares bit for bit between an N-GPU and a 1-GPU run
    assert got[0]["case_records"] == got[1]["case_records"] == [[10.0 + i, float(i), 96.0 * i] for i in range(5)]
    assert len(got[0]["seconds_per_rank"]) == 2 and got[0]["imbalance"] >= 1.0


def test_run_cases_with_several_agents_in_flight():
    """"One or more cases per GPU" (SURVEY 8e): the cases of a rank dealt to several agents, one host thread each -- every case is
    computed once, and the per-case records and the reduced metrics are those of the single-agent run, whatever the interleaving."""
    import threading
    import time as _time
    from vaevar_b200.dist import run_cases
    seen = []

    class Agent(_FakeAgent):
        def one_step_DA(self, gt, xb, yo, H, R, mode="vae4dvar"):
            seed = float(xb.flatten()[0])
            _time.sleep(0.002 * ((int(seed) * 7) % 5))                 # uneven case lengths: the threads interleave
            seen.append((int(seed), threading.get_ident()))
            self.metrics_list["ana_wrmse"].append(torch.full((self.nchannel,), seed))
            self.metrics_list["ana_bias"].append(torch.full((self.nchannel,), -seed))
            self.history.append({"loss": 10.0 + seed, "gmax": 0.5 * seed})
            return xb

    def agents(n):
        out = [Agent() for _ in range(n)]
        for a in out:
            a.history = []
        return out

    mk = lambda i: {k: torch.full((1, 3, 4, 8), float(i)) for k in ("gt", "yo", "H", "R")} | {"xb": torch.full((3, 4, 8), float(i))}
    one = run_cases(agents(1)[0], 11, mk, 0, 1, "cpu")
    seen.clear()
    three = run_cases(agents(3), 11, mk, 0, 1, "cpu")
    assert sorted(c for c, _ in seen) == list(range(11)) and len({t for _, t in seen}) == 3
    assert three["cases_in_flight_per_gpu"] == 3 and one["cases_in_flight_per_gpu"] == 1
    for k in ("n_cases", "mean_J", "mean_gmax", "rms_wrmse", "mean_bias", "case_records"):
        assert one[k] == three[k], k


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
        assert got[r] == serial.buf.tolist()


class _FakeAgent:
    """Host-side stand-in for VaeVar4D: x -> 0.5 x forecast, analysis = background + 1."""
    def __init__(self):
        self.device = torch.device("cpu")
        self.da_win, self.nchannel, self.nlat, self.nlon = 2, 3, 4, 8
        self.metrics_list = {k: [] for k in ("bg_wrmse", "bg_bias", "ana_wrmse", "ana_bias")}

    def integrate(self, x, model=None, step=1, interpolation=False, detach=True):
        return x * (0.5 ** step)

    def one_step_DA(self, gt, xb, yo, H, R, mode="vae4dvar"):
        for k in self.metrics_list:
            self.metrics_list[k].append(torch.full((self.nchannel,), float(len(self.metrics_list[k]))))
        return xb + 1.0


class _Obs:
    def __init__(self):
        self.calls = []

    def window(self, cycle):
        self.calls.append(cycle)
        z = torch.zeros(2, 3, 4, 8)
        return z, z, z + 1, z


def test_cycled_da_resumes_from_its_checkpoint(tmp_path):
    """run_assimilation / save_ckpt / get_current_states / load_eval_ckpts (da_4dvar.py:1314-1342, 683-727): a run cut after
    two cycles and resumed gives the same background and metric history as an uninterrupted one."""
    from vaevar_b200.cycle import CycledDA
    xb0 = torch.ones(3, 4, 8)
    full = CycledDA(_FakeAgent(), _Obs(), xb0, name="full", root=str(tmp_path), n_cycles=4)
    r = full.run_assimilation()
    assert r["cycles"] == 4 and (tmp_path / "full" / "xb.npy").exists()
    part = CycledDA(_FakeAgent(), _Obs(), xb0, name="cut", root=str(tmp_path), n_cycles=2)
    part.run_assimilation()
    obs = _Obs()
    rest = CycledDA(_FakeAgent(), obs, xb0, name="cut", root=str(tmp_path), n_cycles=4)      # picks up xb.npy / current_time.txt
    assert rest.current_cycle == 2 and len(rest.agent.metrics_list["ana_wrmse"]) == 2
    rest.run_assimilation()
    assert obs.calls == [2, 3]
    assert torch.equal(rest.xb, full.xb)
    assert (tmp_path / "cut" / "current_time.txt").read_text() == "4"
    import numpy as np
    assert np.load(tmp_path / "cut" / "ana_wrmse.npy").shape == (4, 3)


def test_obs_interpolater_matches_reference_class(gold):
    """vaevar_b200.da.obs_interpolater against the matrices produced by the reference's own class (da_4dvar.py:62-94; cut out of
    its source and executed by tools/make_golden.py::reference_obs_interp), bit for bit; and the tap table built from it."""
    from vaevar_b200.da import obs_interpolater
    from vaevar_b200.engine import obs_taps
    g = gold("obs_interp.npz")
    o = obs_interpolater(13, 40)
    assert np.array_equal(o.interp.numpy(), g["interp"]) and np.array_equal(o.interp_inv.numpy(), g["interp_inv"])
    assert np.array_equal(o.height_level_new, g["height_level_new"])
    chan, w = obs_taps(g["interp"])
    assert chan.shape == (204, 2) and w.shape == (204, 2)
    dense = np.zeros((204, 69), np.float32)                       # the augmentation of da_4dvar.py:1196-1206 as one matrix
    dense[:4, :4] = np.eye(4)
    for v in range(5):
        dense[4 + 40 * v:4 + 40 * (v + 1), 4 + 13 * v:4 + 13 * (v + 1)] = g["interp"]
    mine = np.zeros_like(dense)
    for a in range(204):
        for j in range(2):
            mine[a, chan[a, j]] += w[a, j]
    assert np.array_equal(mine, dense)


def test_real_simu_observation_source(gold):
    """RealSimuObs against the reference's formulas (da_4dvar.py:745-756, 766-800): augmented truth with the reference's own
    interpolation matrix, yo = aug(gt) * H, R = aug(R_static), quality-control filter."""
    from vaevar_b200.cycle import RealSimuObs, augment_levels
    from vaevar_b200.da import obs_interpolater
    from vaevar_b200.synth import obs_variance
    g = gold("obs_interp.npz")

    class Agent(_FakeAgent):
        def __init__(self):
            super().__init__()
            self.nchannel, self.da_win = 69, 2
            self.obs_interp = obs_interpolater(13, 40)

    a = Agent()
    truth0 = torch.randn(69, 4, 8)
    src = RealSimuObs(a, truth0, obs_frac=0.3, seed=1)
    yo, H, R, gt = src.window(0)
    assert yo.shape == H.shape == R.shape == (2, 204, 4, 8) and gt.shape == (2, 69, 4, 8)
    interp = torch.from_numpy(g["interp"])
    ref = [gt[:, :4]] + [torch.nn.functional.linear(gt[:, 4 + 13 * i:4 + 13 * (i + 1)].transpose(1, 3), interp).transpose(1, 3) for i in range(5)]
    ref = torch.cat(ref, 1)                                                     # the reference's expression, da_4dvar.py:770-776
    torch.testing.assert_close(augment_levels(gt, interp), ref, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(yo, ref * H, rtol=1e-6, atol=1e-6)
    assert set(np.unique(H.numpy())) <= {0.0, 1.0} and 0.2 < float(H.mean()) < 0.4
    var = torch.from_numpy(obs_variance(0.005, 2).astype(np.float32)).reshape(1, 69, 1, 1).expand(2, 69, 4, 8)
    Rref = torch.cat([var[:, :4]] + [torch.nn.functional.linear(var[:, 4 + 13 * i:4 + 13 * (i + 1)].transpose(1, 3), interp).transpose(1, 3)
                                      for i in range(5)], 1)                     # get_R_matrix_from_gt, :745-756
    torch.testing.assert_close(R, Rref, rtol=1e-6, atol=0)
    yo2, H2, _, _ = src.window(0)
    assert torch.equal(H, H2) and torch.equal(yo, yo2)                           # reproducible per cycle
    # quality control (:780-787): a "real" observation 10 sigma off the truth is filtered out, the rest stays
    bad = torch.zeros(2, 204, 4, 8); bad[:, 7] = 10.0
    src_qc = RealSimuObs(a, truth0, obs_frac=0.3, seed=1, filter_coeff=3.0, std_layer_aug=np.ones(204, np.float32),
                         yo_real=lambda cycle: ref + bad)
    _, Hq, _, _ = src_qc.window(0)
    assert float(Hq[:, 7].sum()) == 0.0 and torch.equal(Hq[:, :7], H[:, :7]) and torch.equal(Hq[:, 8:], H[:, 8:])


def test_line_search_cubic_interpolation_matches_torch_and_survives_overflow():
    """The controller's cubic interpolation (host code of libvaevar.so, no device needed) equals torch.optim.lbfgs._cubic_interpolate on
    float32-tensor scalars, and where torch's float32 arithmetic overflows (losses ~1e10 over steps ~1e-11: d1 * d1 = inf, the step becomes
    inf / inf = NaN and poisons z) it returns the finite bisection step instead -- the native-geometry chains hit exactly that."""
    from torch.optim.lbfgs import _cubic_interpolate
    from vaevar_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for _ in range(200):
        x1, x2 = sorted(rng.uniform(1e-4, 2.0, size=2))
        f1, f2 = rng.uniform(-5, 5, size=2)
        g1, g2 = rng.uniform(-10, 10, size=2)
        bounds = bool(rng.integers(0, 2))
        lo, hi = (x2 + 0.01 * (x2 - x1), 10 * x2) if bounds else (0.0, 0.0)
        T = lambda v: torch.tensor(v, dtype=torch.float32)
        ref = _cubic_interpolate(T(x1), float(f1), T(g1), T(x2), float(f2), T(g2), bounds=(T(lo), T(hi)) if bounds else None)
        got = lib.vv_debug_cubic_interpolate(x1, f1, g1, x2, f2, g2, 1, 1, int(bounds), lo, hi)
        assert abs(got - float(ref)) <= 2e-6 * max(1.0, abs(float(ref))), (x1, f1, g1, x2, f2, g2, bounds, got, float(ref))
    # the overflow case: torch yields NaN, the controller the midpoint of the bracket
    x1, x2, f1, f2, g1, g2 = 0.0, 1e-11, 4.1e10, 2.9e10, -3.0e15, -2.5e15
    T = lambda v: torch.tensor(v, dtype=torch.float32)
    ref = _cubic_interpolate(T(x1), f1, T(g1), T(x2), f2, T(g2))
    assert not torch.isfinite(torch.as_tensor(ref)), "torch's interpolation is expected to overflow here"
    got = lib.vv_debug_cubic_interpolate(x1, f1, g1, x2, f2, g2, 1, 1, 0, 0.0, 0.0)
    assert np.isfinite(got) and abs(got - 0.5e-11) <= 1e-17
