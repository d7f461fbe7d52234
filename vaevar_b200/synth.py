"""Seeded synthetic weights and assimilation cases (no datasets, no checkpoints).

Weights: a `state_dict` with the reference's parameter names and shapes
(networks_old/transformer.py, networks_old/utils/swinblock.py)
drawn with numpy's PCG64 so the same seed gives the same bytes on every machine.
Magnitudes follow the reference initialisers: Linear / pos-embed / bias-table
N(0, 0.02) (`trunc_normal_(std=.02)`, transformer.py:381-388, swinblock.py:118),
Conv2d / ConvTranspose2d U(+-1/sqrt(fan_in)) (torch default), LayerNorm 1/0.
`gain` scales the Linear weights (x3-5 makes softmax / GELU / shift mask matter),
`rich=True` also randomises biases and LayerNorm affines.

Cases: the conventions of da_4dvar.py the synthetic generator has to copy --
random column mask shared by all 69 channels and all T (:276-292), noise-free
observations `yo = gt` (:442-450), `R = (obs_std * sigma_c)^2` with the
`modify_tp=2` rescaling (:106-119), Q = 0 (:540-541).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .config import NetConfig, era5_stats


def _block(sd: Dict[str, np.ndarray], rng, pre: str, d: int, heads: int, ws: int, gain: float, rich: bool):
    def lin(name, out_f, in_f, bias=True):
        sd[f"{pre}.{name}.weight"] = (rng.standard_normal((out_f, in_f), dtype=np.float32) * (0.02 * gain))
        if bias:
            sd[f"{pre}.{name}.bias"] = (rng.standard_normal(out_f, dtype=np.float32) * 0.02 if rich
                                        else np.zeros(out_f, np.float32))

    def ln(name, n):
        sd[f"{pre}.{name}.weight"] = (1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32) if rich
                                      else np.ones(n, np.float32))
        sd[f"{pre}.{name}.bias"] = (0.05 * rng.standard_normal(n, dtype=np.float32) if rich
                                    else np.zeros(n, np.float32))

    ln("norm1", d)
    sd[f"{pre}.attn.relative_position_bias_table"] = (
        rng.standard_normal(((2 * ws - 1) ** 2, heads), dtype=np.float32) * (0.02 * (10.0 if rich else 1.0)))
    lin("attn.qkv", 3 * d, d)
    lin("attn.proj", d, d)
    ln("norm2", d)
    lin("mlp.fc1", 4 * d, d)
    lin("mlp.fc2", d, 4 * d)


def make_state_dict(cfg: NetConfig, seed: int = 0, gain: float = 1.0, rich: bool = False) -> Dict[str, np.ndarray]:
    """All learnable tensors of one `LGUnet_all`, float32 numpy, reference key names."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: Dict[str, np.ndarray] = {}
    D, E, ws, G = cfg.enc_dim, cfg.embed_dim, cfg.window_size, cfg.groups
    L0 = cfg.res0[0] * cfg.res0[1]
    L1 = cfg.res1[0] * cfg.res1[1]

    def normal(shape, std=0.02):
        return rng.standard_normal(shape, dtype=np.float32) * np.float32(std)

    def uniform(shape, bound):
        return ((rng.random(shape, dtype=np.float32) * 2.0 - 1.0) * np.float32(bound)).astype(np.float32)

    def ln(pre, n):
        sd[pre + ".weight"] = (1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32) if rich else np.ones(n, np.float32))
        sd[pre + ".bias"] = (0.05 * rng.standard_normal(n, dtype=np.float32) if rich else np.zeros(n, np.float32))

    def lin(pre, out_f, in_f, bias=True):
        sd[pre + ".weight"] = normal((out_f, in_f), 0.02 * gain)
        if bias:
            sd[pre + ".bias"] = normal(out_f) if rich else np.zeros(out_f, np.float32)

    for g in range(G):
        p = f"enc.enc_list.{g}"
        cin = cfg.inchans_list[g]
        sd[p + ".absolute_pos_embed"] = normal((1, L0, D))
        b = 1.0 / np.sqrt(cin * 4)
        sd[p + ".patch_embed.proj.weight"] = uniform((D, cin, 2, 2), b)
        sd[p + ".patch_embed.proj.bias"] = uniform((D,), b)
        for blk in range(cfg.enc_depths[0]):
            _block(sd, rng, f"{p}.layers.0.blocks.{blk}", D, cfg.enc_heads[0], ws, gain, rich)
        lin(p + ".layers.1.downsample.reduction", 2 * D, 4 * D, bias=False)
        ln(p + ".layers.1.downsample.norm", 4 * D)
        for blk in range(cfg.enc_depths[1]):
            _block(sd, rng, f"{p}.layers.1.blocks.{blk}", 2 * D, cfg.enc_heads[1], ws, gain, rich)
        ln(p + ".norm", 2 * D)
    lin("enc.proj", E, 2 * D * G)

    sd["net.pos_embed"] = normal((1, L1, E))
    for l, depth in enumerate(cfg.lg_depths):
        for blk in range(depth):
            _block(sd, rng, f"net.layers.{l}.blocks.{blk}", E, cfg.lg_heads[l], ws, gain, rich)

    lin("dec.proj", 2 * D * G, E)
    for g in range(G):
        p = f"dec.dec_list.{g}"
        cout = cfg.outchans_list[g]
        for blk in range(cfg.enc_depths[1]):
            _block(sd, rng, f"{p}.layers_up.0.blocks.{blk}", 2 * D, cfg.enc_heads[1], ws, gain, rich)
        lin(p + ".layers_up.0.upsample.expand", 4 * D, 2 * D, bias=False)
        ln(p + ".layers_up.0.upsample.norm", D)
        for blk in range(cfg.enc_depths[0]):
            _block(sd, rng, f"{p}.layers_up.1.blocks.{blk}", D, cfg.enc_heads[0], ws, gain, rich)
        lin(p + ".concat_back_dim.0", 2 * D, 4 * D)
        lin(p + ".concat_back_dim.1", D, 2 * D)
        ln(p + ".norm_up", D)
        b = 1.0 / np.sqrt(cout * 4)
        sd[f"dec.final_proj_list.{g}.weight"] = uniform((D, cout, 2, 2), b)
        sd[f"dec.final_proj_list.{g}.bias"] = uniform((cout,), b)
    return sd


def obs_variance(obs_std: float = 0.005, modify_tp: int = 2) -> np.ndarray:
    """Per-channel observation-error variance, float32[69] (da_4dvar.py:106-127)."""
    _, std, _ = era5_stats()
    std32 = std.astype(np.float32)
    var = np.full(69, obs_std, np.float32) ** 2 * std32 ** 2
    if modify_tp == 1:
        var[56:] /= 4
    elif modify_tp in (2, 3, 4):
        var[56:] /= 16
        var[2] /= 16
        if modify_tp == 3:
            var[30:56] /= 16
        if modify_tp == 4:
            var[17:30] /= 4
    return var.astype(np.float32)


def make_case(T: int, nlat: int = 128, nlon: int = 256, obs_frac: float = 0.10, seed: int = 0,
              obs_std: float = 0.005, modify_tp: int = 2, z_std: float = 0.1, latent: int = 32):
    """One synthetic assimilation case.

    Returns a dict of float32 numpy arrays:
      gt (T,69,nlat,nlon) truth, xb (69,..) background, yo (T,69,..) observations,
      H (T,69,..) 0/1 mask, R (T,69,..) obs-error variance, z (1,latent,nlat,nlon).
    """
    mean, std, _ = era5_stats()
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    m = mean.astype(np.float32).reshape(1, 69, 1, 1)
    s = std.astype(np.float32).reshape(1, 69, 1, 1)
    gt = (m + s * rng.standard_normal((T, 69, nlat, nlon), dtype=np.float32)).astype(np.float32)
    xb = (gt[0] + 0.1 * s[0] * rng.standard_normal((69, nlat, nlon), dtype=np.float32)).astype(np.float32)
    n_cols = int(obs_frac * nlat * nlon)
    cols = rng.choice(nlat * nlon, n_cols, replace=False)
    mask = np.zeros(nlat * nlon, np.float32)
    mask[cols] = 1.0
    H = np.broadcast_to(mask.reshape(1, 1, nlat, nlon), (T, 69, nlat, nlon)).copy()
    var = obs_variance(obs_std, modify_tp)
    R = np.broadcast_to(var.reshape(1, 69, 1, 1), (T, 69, nlat, nlon)).copy()
    z = (z_std * rng.standard_normal((1, latent, nlat, nlon), dtype=np.float32)).astype(np.float32)
    return dict(gt=gt, xb=xb, yo=gt.copy(), H=H, R=R, z=z)


def make_real_obs(gt: np.ndarray, interp: np.ndarray, frac: float = 0.02, seed: int = 0, obs_std: float = 0.005, nlev: int = 13):
    """Synthetic stand-in for data_reader.get_real_obs + get_R_matrix_from_gt (da_4dvar.py:745-756, 766-800, obs_type "real_simu"):
    observations live in the AUGMENTED space (4 surface channels + 5 variables x interp.shape[0] pressure levels), each augmented
    channel has its own sparse random mask (a different one per time level), yo = H * aug(gt), R = aug(obs variance).
    Returns dict(yo, H, R) of shape (T, 4 + 5 * dim_out, nlat, nlon), float32."""
    T, _, nlat, nlon = gt.shape
    rng = np.random.Generator(np.random.PCG64(2000 + seed))

    def aug(x):
        parts = [x[:, :4]]
        for i in range(5):
            parts.append(np.einsum("ol,tlhw->tohw", interp.astype(np.float32), x[:, 4 + i * nlev:4 + (i + 1) * nlev]).astype(np.float32))
        return np.concatenate(parts, 1)

    gt_aug = aug(gt)
    H = (rng.random(gt_aug.shape, dtype=np.float32) < frac).astype(np.float32)
    var = obs_variance(obs_std, 2).astype(np.float32).reshape(1, 69, 1, 1)
    R = np.broadcast_to(aug(np.broadcast_to(var, (T, 69, 1, 1))), gt_aug.shape).copy()
    return dict(yo=(gt_aug * H).astype(np.float32), H=H, R=R)


def net1_param_shapes(cfg) -> Dict[str, tuple]:
    """state_dict names and shapes of `LGUnet_all_1` (networks/LGUnet_all.py:743-777), in the reference's registration order per
    module; pinned to the reference module's own state_dict by tests/test_host_logic.py (fixture tests/golden/net1_mid.npz)."""
    sh: Dict[str, tuple] = {}
    G, D, E = cfg.groups, cfg.enc_dim, cfg.embed_dim
    nl = len(cfg.enc_depths)
    h0, w0 = cfg.patches
    kh, kw = cfg.patch_size

    def block(p, d):
        sh[p + ".norm.weight"] = (d,); sh[p + ".norm.bias"] = (d,)
        sh[p + ".attn.qkv.weight"] = (3 * d, d); sh[p + ".attn.qkv.bias"] = (3 * d,)
        sh[p + ".attn.proj.weight"] = (d, d); sh[p + ".attn.proj.bias"] = (d,)
        sh[p + ".norm2.weight"] = (d,); sh[p + ".norm2.bias"] = (d,)
        sh[p + ".mlp.fc1.weight"] = (4 * d, d); sh[p + ".mlp.fc1.bias"] = (4 * d,)
        sh[p + ".mlp.fc2.weight"] = (d, 4 * d); sh[p + ".mlp.fc2.bias"] = (d,)

    for g in range(G):
        p = f"enc.enc_list.{g}"
        sh[p + ".absolute_pos_embed"] = (1, h0 * w0, D)
        sh[p + ".patch_embed.proj.weight"] = (D, cfg.inchans_list[g], kh, kw)
        sh[p + ".patch_embed.proj.bias"] = (D,)
        for l in range(nl):
            d = D << l
            for b in range(cfg.enc_depths[l]):
                block(f"{p}.layers.{l}.blocks.{b}", d)
            if l > 0:
                sh[f"{p}.layers.{l}.downsample.reduction.weight"] = (d, 2 * d)
                sh[f"{p}.layers.{l}.downsample.norm.weight"] = (2 * d,); sh[f"{p}.layers.{l}.downsample.norm.bias"] = (2 * d,)
        sh[p + ".norm.weight"] = (D << (nl - 1),); sh[p + ".norm.bias"] = (D << (nl - 1),)
    top = D << (nl - 1)
    sh["enc.proj.weight"] = (E, G * top); sh["enc.proj.bias"] = (E,)
    ht, wt = cfg.level_grid(nl - 1)
    sh["net.pos_embed"] = (1, ht * wt, E)
    for s, depth in enumerate(cfg.lg_depths):
        for b in range(depth):
            block(f"net.layers.{s}.blocks.{b}", E)
    for g in range(G):
        p = f"dec.dec_list.{g}"
        for inx in range(nl):
            d = D << (nl - 1 - inx)
            for b in range(cfg.enc_depths[nl - 1 - inx]):
                block(f"{p}.layers_up.{inx}.blocks.{b}", d)
            if inx < nl - 1:
                sh[f"{p}.layers_up.{inx}.upsample.expand.weight"] = (2 * d, d)
                sh[f"{p}.layers_up.{inx}.upsample.norm.weight"] = (d // 2,); sh[f"{p}.layers_up.{inx}.upsample.norm.bias"] = (d // 2,)
        for inx in range(nl):
            d = D << (nl - 1 - inx)
            sh[f"{p}.concat_back_dim.{inx}.weight"] = (d, 2 * d); sh[f"{p}.concat_back_dim.{inx}.bias"] = (d,)
        sh[p + ".norm_up.weight"] = (D,); sh[p + ".norm_up.bias"] = (D,)
    for g in range(G):
        sh[f"dec.final_proj_list.{g}.weight"] = (D, cfg.outchans_list[g], kh, kw)
        sh[f"dec.final_proj_list.{g}.bias"] = (cfg.outchans_list[g],)
    sh["dec.proj.weight"] = (G * top, E); sh["dec.proj.bias"] = (G * top,)
    return sh


def make_state_dict_net1(cfg, seed: int = 0, rich: bool = False) -> Dict[str, np.ndarray]:
    """Random-init weights of `LGUnet_all_1` by reference name: Linear N(0, 0.02) with zero bias and LayerNorm (1, 0) as
    `_init_weights` leaves them (networks/LGUnet_all.py:763-770), embeddings N(0, 0.02), convolutions U(+-1/sqrt(fan_in)).
    rich: LayerNorm affines around (1, 0), non-zero biases and 4x larger Linear weights, so that no parameter is silent."""
    rng = np.random.Generator(np.random.PCG64(7000 + seed))
    sd: Dict[str, np.ndarray] = {}
    shapes = net1_param_shapes(cfg)
    for name, shape in shapes.items():
        r = rng.standard_normal(shape, dtype=np.float32)
        conv_w = name[: -len("bias")] + "weight" if name.endswith("bias") else name
        if len(shapes.get(conv_w, ())) == 4:                     # Conv2d / ConvTranspose2d weight and bias: default uniform init
            w = shapes[conv_w]
            fan = w[1] * w[2] * w[3]
            v = (rng.random(shape, dtype=np.float32) * 2 - 1) / np.float32(np.sqrt(fan))
        elif name.endswith("embed"):
            v = 0.02 * r
        elif name.endswith("weight") and len(shape) == 1:        # LayerNorm weight
            v = 1.0 + (0.1 if rich else 0.0) * r
        elif name.endswith("weight"):                            # Linear weight
            v = (0.08 if rich else 0.02) * r
        else:                                                    # Linear / LayerNorm bias
            v = (0.05 if rich else 0.0) * r
        sd[name] = np.ascontiguousarray(v, dtype=np.float32)
    return sd
