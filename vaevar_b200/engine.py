"""Python handle on the CUDA engine (one per process / GPU).  Plumbing only: torch supplies device memory and
streams, every computation happens in libvaevar.so."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .config import NetConfig, era5_stats


def _net_c(cfg: NetConfig, keep_out: int = 0) -> _lib.NetConfigC:
    c = _lib.NetConfigC()
    c.img_h, c.img_w = cfg.img_size
    c.n_groups = cfg.groups
    for i, v in enumerate(cfg.inchans_list):
        c.in_chans[i] = v
    for i, v in enumerate(cfg.outchans_list):
        c.out_chans[i] = v
    c.enc_dim, c.embed_dim, c.window = cfg.enc_dim, cfg.embed_dim, cfg.window_size
    c.enc_depth[0], c.enc_depth[1] = cfg.enc_depths
    c.enc_heads[0], c.enc_heads[1] = cfg.enc_heads
    c.n_lg = len(cfg.lg_depths)
    for i, (d, h) in enumerate(zip(cfg.lg_depths, cfg.lg_heads)):
        c.lg_depth[i], c.lg_heads[i] = d, h
    c.keep_out = keep_out
    return c


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev32(x, device) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=torch.float32).contiguous()


class Engine:
    """cost J(z) and grad_z J of da_4dvar.py:1183-1208 / 1242-1246 on one B200."""

    def __init__(self, dec: NetConfig, flow: Optional[NetConfig] = None, T: int = 1, recompute: bool = False,
                 use_graph: bool = True, device: str = "cuda:0", flow_keep: int = 69, dec_keep: int = 0,
                 forward_fp16: bool = True, ln_fold: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("vaevar_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        torch.cuda.current_stream()          # make sure the primary context exists before the library touches it
        _lib.check(self.lib.vv_set_device(self.device.index or 0))
        self.dec_cfg, self.flow_cfg, self.T = dec, flow, T
        cfg = _lib.ConfigC()
        cfg.dec = _net_c(dec, dec_keep)
        if flow is not None:
            cfg.flow = _net_c(flow, flow_keep)
        cfg.has_flow = int(flow is not None)
        cfg.T, cfg.recompute, cfg.use_graph = T, int(recompute), int(use_graph)
        cfg.forward_fp16 = int(forward_fp16)     # fp16 forward / bf16 gradients (include/vaevar.h: vv_config.forward_fp16)
        cfg.no_ln_fold = int(not ln_fold)        # False: LayerNorm kernels + plain GEMMs (include/vaevar.h: vv_config.no_ln_fold)
        self.ln_fold = bool(ln_fold)
        self._h = C.c_void_p()
        _lib.check(self.lib.vv_engine_create(C.byref(cfg), C.byref(self._h)))
        self.n_state = dec_keep or dec.out_chans
        self.n_latent = dec.in_chans
        self.grid = dec.img_size
        self._keep = []
        mean, std, stdtr = era5_stats()
        if self.n_state == len(mean):
            self.set_constants(mean, std, stdtr)

    # -- weights --------------------------------------------------------------------------------
    def load_state_dict(self, net: int, sd: Dict[str, "np.ndarray | torch.Tensor"]):
        """net 0 = decoder (VAE_lr.dec), 1 = flow model; keys are the reference state_dict names."""
        for k, v in sd.items():
            if "relative_position_index" in k or "attn_mask" in k:
                continue
            t = _dev32(v, self.device)
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self.lib.vv_set_weight(self._h, net, k.encode(), _ptr(t), shape, t.dim()))
        torch.cuda.synchronize()

    def finalize(self):
        _lib.check(self.lib.vv_finalize_weights(self._h))

    def set_constants(self, mean, std, stdtr):
        a = [np.ascontiguousarray(np.asarray(x, np.float32)) for x in (mean, std, stdtr)]
        _lib.check(self.lib.vv_set_constants(self._h, *[x.ctypes.data_as(C.c_void_p) for x in a]))

    def ln_fold_health(self, raise_on_risk: bool = False):
        """Counters of the folded LayerNorms since the last call: rows whose mean drifted > 32 sigma from their stage-input mean
        (`far_mean`) or whose centred values approach the fp16 range (`near_saturation`).  Non-zero = use Engine(ln_fold=False)."""
        c = (C.c_uint32 * 2)()
        _lib.check(self.lib.vv_ln_fold_health(self._h, c))
        out = {"far_mean": int(c[0]), "near_saturation": int(c[1])}
        if raise_on_risk and (out["far_mean"] or out["near_saturation"]):
            raise _lib.VVError(f"folded LayerNorm at risk for these weights / inputs ({out}): construct the engine with ln_fold=False")
        return out

    # -- case -----------------------------------------------------------------------------------
    def set_case(self, xb, yo, H, R, obs_coeff: float = 1.0):
        xb, yo, H, R = (_dev32(x, self.device) for x in (xb, yo, H, R))
        _lib.check(self.lib.vv_set_case(self._h, _ptr(xb), _ptr(yo), _ptr(H), _ptr(R), float(obs_coeff), _stream()))

    def set_case_native(self, xb, yo, H, R, obs_coeff: float = 1.0):
        """The closure on the reference's real geometry: xb (C,Hh,Wh), yo / H / R (T,C,Hh,Wh) on an analysis grid finer than the
        network grid (decoder_hr, integrate(..., interpolation=True); nf_model/vae.py:87-90, da_4dvar.py:666-681, 1185-1208)."""
        xb, yo, H, R = (_dev32(x, self.device) for x in (xb, yo, H, R))
        self.native_grid = tuple(int(v) for v in xb.shape[-2:])
        _lib.check(self.lib.vv_set_case_native(self._h, _ptr(xb), _ptr(yo), _ptr(H), _ptr(R), self.native_grid[0], self.native_grid[1],
                                               float(obs_coeff), _stream()))

    def set_case_real_obs(self, xb, yo, H, R, interp, obs_coeff: float = 1.0, n_surface: int = 4, n_vars: int = 5):
        """The real-observation branch (da_4dvar.py:1196-1206): yo / H / R (T, n_surface + n_vars * dim_out, Hh, Wh) in the augmented
        space, `interp` = obs_interpolater.interp (dim_out, nlev) (da_4dvar.py:62-82).  The grid may be the network grid or finer."""
        interp = np.asarray(interp.detach().cpu() if isinstance(interp, torch.Tensor) else interp, np.float32)
        tap_chan, tap_w = obs_taps(interp, n_surface, n_vars)
        xb, yo, H, R = (_dev32(x, self.device) for x in (xb, yo, H, R))
        if yo.shape[1] != tap_chan.shape[0]:
            raise ValueError(f"yo has {yo.shape[1]} channels, the operator produces {tap_chan.shape[0]}")
        self.native_grid = tuple(int(v) for v in xb.shape[-2:])
        _lib.check(self.lib.vv_set_case_obsop(self._h, _ptr(xb), _ptr(yo), _ptr(H), _ptr(R), self.native_grid[0], self.native_grid[1],
                                              tap_chan.shape[0], tap_chan.shape[1], tap_chan.ctypes.data_as(C.c_void_p),
                                              tap_w.ctypes.data_as(C.c_void_p), float(obs_coeff), _stream()))

    def decode_native(self, z: torch.Tensor) -> torch.Tensor:
        """(decoder_hr(z) stdTr) sigma + xb on the analysis grid (da_4dvar.py:1257-1259, 1301-1306)."""
        out = torch.empty(self.n_state, *self.native_grid, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.vv_decode_native(self._h, _ptr(z.contiguous()), _ptr(out), _stream()))
        return out

    @property
    def n_obs(self) -> int:
        n = C.c_int64()
        _lib.check(self.lib.vv_num_obs(self._h, C.byref(n)))
        return n.value

    # -- evaluation -----------------------------------------------------------------------------
    def cost_grad(self, z: torch.Tensor, J_out: Optional[torch.Tensor] = None, grad_out: Optional[torch.Tensor] = None):
        """One closure(): returns (J[3] float64 device tensor = {J, J_reg, J_obs}, grad like z). Asynchronous."""
        z = z.contiguous()
        J = torch.empty(3, dtype=torch.float64, device=self.device) if J_out is None else J_out
        g = torch.empty_like(z) if grad_out is None else grad_out
        _lib.check(self.lib.vv_cost_grad(self._h, _ptr(z), _ptr(J), _ptr(g), _stream()))
        return J, g

    def cost(self, z: torch.Tensor):
        J = torch.empty(3, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.vv_cost(self._h, _ptr(z.contiguous()), _ptr(J), _stream()))
        return J

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        out = torch.empty(self.n_state, *self.grid, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.vv_decode(self._h, _ptr(z.contiguous()), _ptr(out), _stream()))
        return out

    def metrics(self, x_phys: torch.Tensor, gt_phys: torch.Tensor):
        """(Metrics.WRMSE, Metrics.Bias) per channel of two PHYSICAL (C,H,W) fields, as da_4dvar.py:1260-1264 computes them
        (utils/metrics.py:526-544, 473-474): one fused device pass, float64 device tensors of length C."""
        out = torch.empty(2 * self.n_state, dtype=torch.float64, device=self.device)
        if x_phys.shape != gt_phys.shape or x_phys.shape[0] != self.n_state:
            raise ValueError("metrics takes two (C,H,W) fields of the same grid")
        _lib.check(self.lib.vv_metrics_grid(self._h, _ptr(x_phys.contiguous()), _ptr(gt_phys.contiguous()), int(x_phys.shape[-2]),
                                            int(x_phys.shape[-1]), _ptr(out), _stream()))
        return out[: self.n_state], out[self.n_state:]

    def integrate(self, x: torch.Tensor, steps: int = 1) -> torch.Tensor:
        out = torch.empty_like(x)
        _lib.check(self.lib.vv_integrate(self._h, _ptr(x.contiguous()), _ptr(out), steps, _stream()))
        return out

    def net_forward(self, net: int, x: torch.Tensor) -> torch.Tensor:
        cfg = self.dec_cfg if net == 0 else self.flow_cfg
        keep = self.n_state
        out = torch.empty(keep, *cfg.img_size, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.vv_net_forward(self._h, net, _ptr(x.contiguous()), _ptr(out), _stream()))
        return out

    def net_vjp(self, net: int, x: torch.Tensor, dout: torch.Tensor) -> torch.Tensor:
        din = torch.empty_like(x)
        _lib.check(self.lib.vv_net_vjp(self._h, net, _ptr(x.contiguous()), _ptr(dout.contiguous()), _ptr(din), _stream()))
        return din

    def obs_term(self, xn: torch.Tensor, want_grad: bool = True):
        J = torch.empty(1, dtype=torch.float64, device=self.device)
        g = torch.empty_like(xn) if want_grad else None
        _lib.check(self.lib.vv_test_obs(self._h, _ptr(xn.contiguous()), _ptr(J), _ptr(g), _stream()))
        return J, g

    def profile_ops(self, app: int, bwd: bool, reps: int = 20, flush_l2: bool = False):
        """CUDA-event time of every launch of one application plan: list of dicts.  flush_l2=False: launches back to back (steady
        state, warm L2); True: every timed launch starts on a cold L2 (what an HBM roofline fraction has to be measured on)."""
        cap = 4096
        ms = (C.c_float * cap)(); kind = (C.c_int * cap)(); fl = (C.c_double * cap)(); mnk = (C.c_int * (4 * cap))()
        n = self.lib.vv_profile_ops(self._h, app, int(bwd), reps, ms, kind, fl, mnk, cap, int(flush_l2))
        if n < 0:
            _lib.check(n)
        names = ["gemm", "ln_fwd", "ln_bwd", "attn_fwd", "attn_bwd", "p2t", "t2p", "rope", "sd_attn", "patch32", "convt32", "mlp_fwd", "mlp_bwd", "lin_fwd"]
        return [dict(kind=names[kind[i]], ms=ms[i], flop=fl[i], shape=tuple(mnk[4 * i:4 * i + 4])) for i in range(min(n, cap))]

    @property
    def last_launch_count(self) -> int:
        return self.lib.vv_last_launch_count(self._h)

    def close(self):
        if self._h:
            self.lib.vv_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LBFGS:
    """torch.optim.LBFGS(history_size, max_iter, line_search_fn='strong_wolfe') bound to an Engine (da_4dvar.py:1240)."""

    def __init__(self, engine: Optional[Engine], history_size: int = 10, max_iter: int = 10, testfn_n: int = 0,
                 f_noise_rel: Optional[float] = None):
        self.engine = engine
        self.lib = _lib.load()
        self._h = C.c_void_p()
        if engine is None:      # analytic pairwise-Rosenbrock closure on the device (controller tests)
            _lib.check(self.lib.vv_lbfgs_create_testfn(testfn_n, history_size, max_iter, C.byref(self._h)))
        else:
            _lib.check(self.lib.vv_lbfgs_create(engine._h, history_size, max_iter, C.byref(self._h)))
        if f_noise_rel is not None:     # line-search tolerance to rounding noise in J (include/vaevar.h: vv_lbfgs_set_noise)
            _lib.check(self.lib.vv_lbfgs_set_noise(self._h, float(f_noise_rel)))

    def history(self):
        n = self.lib.vv_lbfgs_history(self._h, None, 0)
        buf = (C.c_double * max(n, 1))()
        self.lib.vv_lbfgs_history(self._h, buf, n)
        return [buf[i] for i in range(n)]

    def steps(self):
        """Trial step length of every closure evaluation (aligned with history(); 0 = evaluation opening a step())."""
        n = self.lib.vv_lbfgs_steps(self._h, None, 0)
        buf = (C.c_double * max(n, 1))()
        self.lib.vv_lbfgs_steps(self._h, buf, n)
        return [buf[i] for i in range(n)]

    def step(self, z: torch.Tensor):
        info = (C.c_double * 8)()
        _lib.check(self.lib.vv_lbfgs_step(self._h, _ptr(z), info, _stream()))
        return dict(loss0=info[0], loss=info[1], n_evals=int(info[2]), n_iter=int(info[3]), t=info[4], gmax=info[5],
                    func_evals=int(info[6]), skipped_evals=int(info[7]))

    def reset(self):
        """State of a freshly constructed optimiser (the reference builds one per cycle), device vectors kept."""
        _lib.check(self.lib.vv_lbfgs_reset(self._h))

    def last_cost(self):
        """(J, J_reg, J_obs) at the point the last step() left z on (cal_loss(z) without another network sweep)."""
        out = (C.c_double * 3)()
        _lib.check(self.lib.vv_lbfgs_last_cost(self._h, out))
        return torch.tensor([out[0], out[1], out[2]], dtype=torch.float64)

    def set_reuse(self, on: bool):
        """Skip the closure() a step() opens with when z is unchanged since the previous step (default on)."""
        _lib.check(self.lib.vv_lbfgs_set_reuse(self._h, int(on)))

    def close(self):
        if self._h:
            self.lib.vv_lbfgs_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def obs_taps(interp: np.ndarray, n_surface: int = 4, n_vars: int = 5):
    """(tap_chan int32 [A, K], tap_w float32 [A, K]) of the augmentation at da_4dvar.py:1196-1206: the surface channels pass through,
    each upper-air variable's nlev model levels go through `interp` (dim_out, nlev); K = most non-zeros in a row of interp."""
    dim_out, nlev = interp.shape
    K = max(1, int(np.count_nonzero(interp, axis=1).max()))
    A = n_surface + n_vars * dim_out
    chan, w = np.zeros((A, K), np.int32), np.zeros((A, K), np.float32)
    for a in range(n_surface):
        chan[a, :], w[a, 0] = a, 1.0
    for v in range(n_vars):
        for o in range(dim_out):
            nz = np.flatnonzero(interp[o])
            a = n_surface + v * dim_out + o
            chan[a, :] = n_surface + v * nlev + (nz[0] if len(nz) else 0)
            for j, l in enumerate(nz):
                chan[a, j], w[a, j] = n_surface + v * nlev + l, interp[o, l]
    return np.ascontiguousarray(chan), np.ascontiguousarray(w)


def compact_mask(H: torch.Tensor, yo: torch.Tensor, R: torch.Tensor):
    """(idx int32, y, 1/R) of the non-zeros of the dense mask, ascending flat order (== torch.nonzero)."""
    lib = _lib.load()
    n = H.numel()
    idx = torch.empty(n, dtype=torch.int32, device=H.device)
    y = torch.empty(n, dtype=torch.float32, device=H.device)
    ri = torch.empty(n, dtype=torch.float32, device=H.device)
    cnt = C.c_int64()
    _lib.check(lib.vv_compact_mask(_ptr(H.contiguous()), _ptr(yo.contiguous()), _ptr(R.contiguous()), n, _ptr(idx), _ptr(y),
                                   _ptr(ri), C.byref(cnt), _stream()))
    k = cnt.value
    return idx[:k], y[:k], ri[:k]
