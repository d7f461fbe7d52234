"""Generate tests/golden/*.npz by running the REAL reference modules on CPU.

Runs only in the build container (needs /root/reference).  Imports the reference's
own `networks_old.transformer.LGUnet_all`, `nf_model.vae.VAE_lr` and
`utils.metrics.Metrics` unmodified through the 3-module import shim of SURVEY.md
appendix C (timm / fairscale / turtle stubs), loads the repo's seeded synthetic
weights into them by name (strict), and records outputs the CPU oracle
(`oracle/`) and the CUDA engine are then checked against.  Nothing here is used
at run time; the fixtures it writes are committed.

    python tools/make_golden.py [--full]     # --full adds the 128x256 fixtures (minutes)
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import pathlib
import sys
import time
import types

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF = "/root/reference"
GOLD = ROOT / "tests" / "golden"


def install_shim():
    """Stub the three absent third-party modules before importing the reference."""
    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def mk(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mk("timm"); mk("timm.models")
    mk("timm.models.layers", DropPath=DropPath, to_2tuple=to_2tuple, trunc_normal_=torch.nn.init.trunc_normal_)
    mk("fairscale"); mk("fairscale.nn"); mk("fairscale.nn.checkpoint")
    mk("fairscale.nn.checkpoint.checkpoint_activations", checkpoint_wrapper=lambda m, **kw: m)
    mk("turtle", forward=None)
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    os.chdir(REF)


def ref_net(cfg, sd_np):
    from networks_old.transformer import LGUnet_all
    with contextlib.redirect_stdout(io.StringIO()):
        net = LGUnet_all(**cfg.to_reference_kwargs()).eval()
    want = {k for k, _ in net.named_parameters()}
    assert want == set(sd_np), (sorted(want ^ set(sd_np))[:10])
    missing, unexpected = net.load_state_dict({k: torch.from_numpy(v) for k, v in sd_np.items()}, strict=False)
    assert not unexpected and all(("relative_position_index" in m or "attn_mask" in m) for m in missing), (missing[:5], unexpected[:5])
    return net


class RefNets:
    def __init__(self, dec, flow):
        self._dec, self._flow = dec, flow

    def decode(self, z):
        return self._dec(z)

    def flow(self, x):
        return self._flow(x)


def sample_idx(n, k=4096, seed=7):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.sort(rng.choice(n, min(k, n), replace=False))


def golden_net(tag, cfg, seed, gain, rich):
    from vaevar_b200.synth import make_state_dict
    sd = make_state_dict(cfg, seed=seed, gain=gain, rich=rich)
    net = ref_net(cfg, sd)
    rng = np.random.Generator(np.random.PCG64(seed + 77))
    x = rng.standard_normal((1, cfg.in_chans, *cfg.img_size), dtype=np.float32)
    dy = rng.standard_normal((1, cfg.out_chans, *cfg.img_size), dtype=np.float32)
    xt = torch.from_numpy(x).requires_grad_(True)
    y = net(xt)
    (y * torch.from_numpy(dy)).sum().backward()
    y = y.detach().numpy(); dx = xt.grad.numpy()
    iy, ix = sample_idx(y.size), sample_idx(dx.size, seed=8)
    np.savez_compressed(GOLD / f"net_{tag}.npz", seed=seed, gain=gain, rich=rich,
                        y_idx=iy, y_val=y.ravel()[iy], y_sum=np.float64(y.astype(np.float64).sum()),
                        y_abs=np.float64(np.abs(y.astype(np.float64)).sum()),
                        dx_idx=ix, dx_val=dx.ravel()[ix], dx_norm=np.float64(np.linalg.norm(dx.astype(np.float64))))
    print(f"net_{tag}: |y|_1={np.abs(y).sum():.6g} |dx|_2={np.linalg.norm(dx):.6g}")
    return sd, net


def golden_cost(tag, cfg_dec, cfg_flow, T, obs_frac, seed, gain, rich, lbfgs_iters=0, nit4=False, nit_track=0):
    from oracle.cost import Case, cost_and_grad, one_step_da
    from vaevar_b200.synth import make_case, make_state_dict
    sd_d = make_state_dict(cfg_dec, seed=seed, gain=gain, rich=rich)
    dec = ref_net(cfg_dec, sd_d)
    flow = None
    if T > 1:
        sd_f = make_state_dict(cfg_flow, seed=seed + 1, gain=gain, rich=rich)
        flow = ref_net(cfg_flow, sd_f)
    nets = RefNets(dec, flow)
    case = make_case(T, *cfg_dec.img_size, obs_frac=obs_frac, seed=seed)
    c = Case(case)
    t0 = time.time()
    J, Jr, Jo, g = cost_and_grad(case["z"], c, nets)
    dt = time.time() - t0
    ig = sample_idx(g.size, 8192, seed=9)
    out = dict(seed=seed, gain=gain, rich=rich, T=T, obs_frac=obs_frac, J=J, J_reg=Jr, J_obs=Jo,
               g_idx=ig, g_val=g.ravel()[ig], g_norm=np.float64(np.linalg.norm(g.astype(np.float64))),
               n_obs=np.int64(case["H"].sum()), seconds=dt, threads=torch.get_num_threads())
    if g.size <= 1 << 17:
        out["g_full"] = g
    if lbfgs_iters:
        r = one_step_da(c, nets, nit=1, max_iter=lbfgs_iters)
        out.update(bg_wrmse=r["bg_wrmse"], ana_wrmse=r["ana_wrmse"], bg_bias=r["bg_bias"], ana_bias=r["ana_bias"],
                   J_history=r["J_history"], n_evals=r["n_evals"])
    if nit4:   # the shipped script's Nit=4 outer steps (da_4dvar_script.sh:14): a converged analysis
        r = one_step_da(c, nets, nit=4, max_iter=lbfgs_iters)
        out.update(ana_wrmse_nit4=r["ana_wrmse"], ana_bias_nit4=r["ana_bias"], J_history_nit4=r["J_history"], n_evals_nit4=r["n_evals"])
    if nit_track:  # one Nit-step run (da_4dvar_script.sh:14) recording the analysis WRMSE / Bias after every outer step
        tr_w, tr_b = [], []
        r = one_step_da(c, nets, nit=nit_track, max_iter=10, log=lambda kk, w, b: (tr_w.append(w.numpy().copy()), tr_b.append(b.numpy().copy())))
        out.update(wrmse_per_outer=np.stack(tr_w), bias_per_outer=np.stack(tr_b), J_history_nit4=r["J_history"], n_evals_nit4=r["n_evals"],
                   ana_wrmse_nit4=r["ana_wrmse"], bg_wrmse=r["bg_wrmse"])
    np.savez_compressed(GOLD / f"cost_{tag}.npz", **out)
    print(f"cost_{tag}: J={J:.8g} J_reg={Jr:.6g} J_obs={Jo:.8g} |g|={out['g_norm']:.6g} ({dt:.1f}s)", flush=True)


def golden_cost_native(tag, cfg_dec, cfg_flow, T, hr, obs_frac, seed, gain, rich, lbfgs_iters=10):
    """The reference's real geometry in miniature: analysis grid `hr` finer than the network grid by the same non-integer ratios
    as 721x1440 / 128x256, so decoder_hr (vae.py:90) and integrate(x, flow, 1, True, False) (da_4dvar.py:1191, 670-679) resample."""
    from oracle.cost import Case, cost_and_grad, one_step_da
    from vaevar_b200.synth import make_case, make_state_dict
    dec = ref_net(cfg_dec, make_state_dict(cfg_dec, seed=seed, gain=gain, rich=rich))
    flow = ref_net(cfg_flow, make_state_dict(cfg_flow, seed=seed + 1, gain=gain, rich=rich))
    nets = RefNets(dec, flow)
    case = make_case(T, *hr, obs_frac=obs_frac, seed=seed)
    z = make_case(1, *cfg_dec.img_size, obs_frac=obs_frac, seed=seed)["z"]
    c = Case(case, lr=cfg_dec.img_size)
    J, Jr, Jo, g = cost_and_grad(z, c, nets)
    r = one_step_da(c, nets, nit=1, max_iter=lbfgs_iters)
    r4 = one_step_da(c, nets, nit=4, max_iter=lbfgs_iters)     # the shipped script's Nit=4 (da_4dvar_script.sh:14)
    ia = sample_idx(r["xa"].size, 8192, seed=10)
    np.savez_compressed(GOLD / f"cost_{tag}.npz", ana_wrmse_nit4=r4["ana_wrmse"], J_history_nit4=r4["J_history"], n_evals_nit4=r4["n_evals"],
                        seed=seed, gain=gain, rich=rich, T=T, hr=np.array(hr), obs_frac=obs_frac, J=J, J_reg=Jr,
                        J_obs=Jo, g_full=g, g_norm=np.float64(np.linalg.norm(g.astype(np.float64))), n_obs=np.int64(case["H"].sum()),
                        bg_wrmse=r["bg_wrmse"], ana_wrmse=r["ana_wrmse"], bg_bias=r["bg_bias"], ana_bias=r["ana_bias"],
                        J_history=r["J_history"], n_evals=r["n_evals"], xa_idx=ia, xa_val=r["xa"].ravel()[ia])
    print(f"cost_{tag}: J={J:.8g} J_reg={Jr:.6g} J_obs={Jo:.8g} |g|={np.linalg.norm(g):.6g} evals={r['n_evals']}", flush=True)


def reference_obs_interp():
    """The reference's own obs_interpolater (da_4dvar.py:62-94) - the file cannot be imported (petrel_client, torch_harmonics at
    :18-25), so the class is cut out of its source with ast and executed with `.cuda()` removed (no GPU in the build container)."""
    import ast
    src = open(f"{REF}/da_4dvar.py").read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "obs_interpolater")
    code = ast.get_source_segment(src, node).replace(".cuda()", "")
    ns = {"np": np, "torch": torch}
    exec(code, ns)
    return ns["obs_interpolater"](13, 40)


def golden_obs_interp():
    oi = reference_obs_interp()
    np.savez_compressed(GOLD / "obs_interp.npz", interp=oi.interp.numpy(), interp_inv=oi.interp_inv.numpy(),
                        height_level=np.asarray(oi.height_level), height_level_new=np.asarray(oi.height_level_new))
    print("obs_interp", oi.interp.shape, oi.interp_inv.shape)


def golden_real_obs(tag, cfg_dec, cfg_flow, T, hr, seed):
    """The real-observation branch of the loss (da_4dvar.py:1196-1206): 204-channel yo / H / R on the analysis grid, the level
    interpolation from the reference's own obs_interpolater, network modules from the reference."""
    from oracle.cost import Case, cost_and_grad, obs_interp_matrix
    from vaevar_b200.synth import make_case, make_real_obs, make_state_dict
    oi = reference_obs_interp()
    interp = oi.interp.numpy()
    assert np.array_equal(interp, obs_interp_matrix(13, 40)), "oracle restatement of get_interp differs from the reference"
    dec = ref_net(cfg_dec, make_state_dict(cfg_dec, seed=seed))
    flow = ref_net(cfg_flow, make_state_dict(cfg_flow, seed=seed + 1))
    case = make_case(T, *hr, obs_frac=0.10, seed=seed)
    case.update(make_real_obs(case["gt"], interp, seed=seed))
    z = make_case(1, *cfg_dec.img_size, obs_frac=0.1, seed=seed)["z"]
    c = Case(case, lr=None if tuple(hr) == tuple(cfg_dec.img_size) else cfg_dec.img_size, interp=interp)
    J, Jr, Jo, g = cost_and_grad(z, c, RefNets(dec, flow))
    np.savez_compressed(GOLD / f"cost_{tag}.npz", seed=seed, T=T, hr=np.array(hr), interp=interp, J=J, J_reg=Jr, J_obs=Jo, g_full=g,
                        g_norm=np.float64(np.linalg.norm(g.astype(np.float64))), n_obs=np.int64(case["H"].sum()),
                        height_level_new=np.asarray(oi.height_level_new))
    print(f"cost_{tag}: J={J:.8g} J_obs={Jo:.8g} |g|={np.linalg.norm(g):.6g} n_obs={int(case['H'].sum())}", flush=True)


def golden_lgunet1():
    """The forecast network LGUnet_all_1 (networks/LGUnet_all.py:743-777) at a small configuration with the real patch (3, 2) /
    stride 2 / window (6, 12) geometry, run by the reference module itself; weights are synthesised by name (oracle.lgunet1)."""
    from networks.LGUnet_all import LGUnet_all_1
    from oracle.lgunet1 import NET1_SMALL, synth_state_dict
    cfg = NET1_SMALL
    with contextlib.redirect_stdout(io.StringIO()):
        net = LGUnet_all_1(**cfg.to_reference_kwargs()).eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth_state_dict(shapes, seed=3), strict=True)
    rng = np.random.Generator(np.random.PCG64(123))
    x = rng.standard_normal((1, 69, *cfg.img_size), dtype=np.float32)
    with torch.no_grad():
        y = net(torch.from_numpy(x)).numpy()
    iy = sample_idx(y.size, 8192, seed=12)
    np.savez_compressed(GOLD / "net1_small.npz", names=np.array(list(shapes)), shapes=np.array([str(list(s)) for s in shapes.values()]),
                        seed=3, x_seed=123, y_idx=iy, y_val=y.ravel()[iy], y_abs=np.float64(np.abs(y.astype(np.float64)).sum()),
                        y_shape=np.array(y.shape))
    print(f"net1_small: {len(shapes)} tensors, {sum(int(np.prod(s)) for s in shapes.values())} params, |y|_1={np.abs(y).sum():.6g}")


def golden_lgunet1_mid():
    """LGUnet_all_1 at a mid-size configuration the CUDA path supports (enc_dim 96, head widths 32 / 32 / 64 / 192, three tower levels,
    6 x 12 windows with 2 x 2 or more windows per level, a 288-token whole-grid trunk stage), run by the reference module itself on
    weights of vaevar_b200.synth.make_state_dict_net1 (rich: no silent parameter).  Stores sampled outputs of the first 69 channels
    (what `model(x)[:, :69]` keeps, da_4dvar.py:674) and of the whole output."""
    from networks.LGUnet_all import LGUnet_all_1
    from vaevar_b200.config import FORECAST_MID
    from vaevar_b200.synth import make_state_dict_net1
    cfg = FORECAST_MID
    with contextlib.redirect_stdout(io.StringIO()):
        net = LGUnet_all_1(**cfg.to_reference_kwargs()).eval()
    sd = make_state_dict_net1(cfg, seed=11, rich=True)
    ref_shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert list(ref_shapes.items()) == [(k, tuple(v.shape)) for k, v in sd.items()], "net1_param_shapes disagrees with the reference module"
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    rng = np.random.Generator(np.random.PCG64(321))
    x = rng.standard_normal((1, 69, *cfg.img_size), dtype=np.float32)
    t0 = time.time()
    with torch.no_grad():
        y = net(torch.from_numpy(x)).numpy()
    y69 = np.ascontiguousarray(y[0, :69])
    iy = sample_idx(y69.size, 16384, seed=13)
    np.savez_compressed(GOLD / "net1_mid.npz", names=np.array(list(ref_shapes)), shapes=np.array([str(list(s)) for s in ref_shapes.values()]),
                        seed=11, x_seed=321, y_idx=iy, y_val=y69.ravel()[iy], y_abs=np.float64(np.abs(y69.astype(np.float64)).sum()),
                        y_max=np.float64(np.abs(y69).max()), y_rms=np.float64(np.sqrt((y69.astype(np.float64) ** 2).mean())),
                        y_chan_rms=np.sqrt((y69.astype(np.float64) ** 2).mean(axis=(1, 2))), y_shape=np.array(y.shape))
    print(f"net1_mid: {len(sd)} tensors, {sum(v.size for v in sd.values())} params, |y|_1={np.abs(y69).sum():.6g} rms={np.sqrt((y69 ** 2).mean()):.4g} "
          f"max={np.abs(y69).max():.4g} ({time.time() - t0:.1f} s)")


def golden_metrics():
    from utils.metrics import Metrics
    rng = np.random.Generator(np.random.PCG64(5))
    from vaevar_b200.config import era5_stats
    std = torch.from_numpy(era5_stats()[1])
    pred = torch.from_numpy(rng.standard_normal((1, 69, 32, 64), dtype=np.float32))
    gt = torch.from_numpy(rng.standard_normal((1, 69, 32, 64), dtype=np.float32))
    m = Metrics()
    np.savez_compressed(GOLD / "metrics.npz", pred=pred.numpy(), gt=gt.numpy(),
                        wrmse=m.WRMSE(pred, gt, None, None, std).numpy(),
                        bias=m.Bias(pred, gt, None, None, std).numpy())
    print("metrics ok")


def golden_vae_surface():
    """VAE_lr('parameters0_old') key names / shapes and the decoder channel shuffle (vae.py:54-90)."""
    from nf_model.vae import VAE_lr
    with contextlib.redirect_stdout(io.StringIO()):
        vae = VAE_lr("parameters0_old")
    names = sorted(f"{k}:{tuple(v.shape)}" for k, v in vae.dec.state_dict().items())
    (GOLD / "decoder_state_dict_keys.txt").write_text("\n".join(names) + "\n")
    print("VAE_lr decoder tensors:", len(names))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    install_shim()
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    from vaevar_b200.config import DECODER_FULL, ENCODER_FULL, FLOW_FULL, small
    ds, fs = small(DECODER_FULL), small(FLOW_FULL)
    jobs = {
        "metrics": golden_metrics,
        "vae": golden_vae_surface,
        "net_small_dec": lambda: golden_net("small_dec", ds, 0, 1.0, False),
        "net_small_flow_rich": lambda: golden_net("small_flow_rich", fs, 1, 3.0, True),
        "net_small_enc": lambda: golden_net("small_enc", small(ENCODER_FULL), 5, 2.0, True),      # VAE_lr.enc (nf_model/vae.py:64, 72-76)
        "cost_small_T1": lambda: golden_cost("small_T1", ds, fs, 1, 0.10, 0, 1.0, False, lbfgs_iters=10, nit4=True),
        "cost_small_T3_rich": lambda: golden_cost("small_T3_rich", ds, fs, 3, 0.10, 2, 3.0, True, lbfgs_iters=10, nit4=True),
        "cost_native_T3_rich": lambda: golden_cost_native("native_T3_rich", ds, fs, 3, (181, 360), 0.10, 4, 3.0, True),
        "obs_interp": golden_obs_interp,
        "net1_small": golden_lgunet1,
        "net1_mid": golden_lgunet1_mid,
        "cost_realobs_T2": lambda: golden_real_obs("realobs_T2", ds, fs, 2, ds.img_size, 5),
        "cost_realobs_native_T2": lambda: golden_real_obs("realobs_native_T2", ds, fs, 2, (181, 360), 6),
        "cost_native_T3_plain": lambda: golden_cost_native("native_T3_plain", ds, fs, 3, (181, 360), 0.10, 0, 1.0, False),
    }
    if a.full:
        jobs.update({
            "net_full_dec": lambda: golden_net("full_dec", DECODER_FULL, 0, 1.0, False),
            "net_full_flow_rich": lambda: golden_net("full_flow_rich", FLOW_FULL, 1, 3.0, True),
            "cost_full_T1": lambda: golden_cost("full_T1", DECODER_FULL, FLOW_FULL, 1, 0.10, 0, 1.0, False, lbfgs_iters=10, nit4=True),
            "cost_full_T6": lambda: golden_cost("full_T6", DECODER_FULL, FLOW_FULL, 6, 0.10, 0, 1.0, False),
            "cost_full_T2_rich": lambda: golden_cost("full_T2_rich", DECODER_FULL, FLOW_FULL, 2, 0.10, 3, 3.0, True),
            # BASELINE.json configs[2]: 12-step window, 5 % observations (one closure; the engine runs it with adjoint recompute)
            "cost_full_T12_obs05": lambda: golden_cost("full_T12_obs05", DECODER_FULL, FLOW_FULL, 12, 0.05, 7, 1.0, False),
            # the headline config with the shipped script's optimisation: Nit=4 x LBFGS.step(max_iter=10), WRMSE after every outer step
            "cost_full_T6_nit4": lambda: golden_cost("full_T6_nit4", DECODER_FULL, FLOW_FULL, 6, 0.10, 0, 1.0, False, nit_track=4),
        })
    for name, fn in jobs.items():
        if a.only and a.only not in name:
            continue
        fn()
