"""ORACLE (test infrastructure, never on the product path).

CPU fp32 restatement, in plain functional PyTorch, of one application of the
reference's U-shaped Swin network `LGUnet_all`:
    networks_old/transformer.py:747-752  (LGUnet_all.forward)
    networks_old/utils/swinblock.py:265-309 (SwinTransformerBlock.forward)
It consumes a `state_dict` with the reference's own key names, so the same
weights drive the reference modules, this oracle and the CUDA engine.
Gradients come from torch.autograd exactly as in the reference
(da_4dvar.py:1242-1246).

Pinned against the real reference modules imported from /root/reference by
tools/make_golden.py -> tests/golden/*.npz (checked in tests/test_oracle_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from vaevar_b200.config import NetConfig

Tensor = torch.Tensor


def rel_pos_index(ws: int) -> Tensor:
    """(ws*ws, ws*ws) index into the (2ws-1)^2 bias table; swinblock.py:92-103."""
    c = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = c[:, :, None] - c[:, None, :] + (ws - 1)
    return rel[0] * (2 * ws - 1) + rel[1]


def shift_mask(H: int, W: int, ws: int, shift: int) -> Tensor:
    """(nW, ws*ws, ws*ws) additive mask, 0 / -100, latitude direction only.

    swinblock.py:236-260: the last `w_slices` entry is slice(0, None), so every
    longitude gets the label of its latitude band (longitude wraps periodically);
    bands are rows [0,H-ws), [H-ws,H-shift), [H-shift,H) of the rolled frame.
    """
    band = torch.zeros(H, W)
    band[H - ws:H - shift] = 1.0
    band[H - shift:] = 2.0
    win = band.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def swin_block(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int, ws: int, shift: int) -> Tensor:
    """x (B,H,W,C) -> (B,H,W,C); swinblock.py:265-309 with WindowAttention :133-172, Mlp :23-29."""
    B, H, W, C = x.shape
    hd = C // heads
    h = F.layer_norm(x, (C,), sd[pre + ".norm1.weight"], sd[pre + ".norm1.bias"], 1e-5)
    if shift:
        h = torch.roll(h, (-shift, -shift), (1, 2))
    nWh, nWw = H // ws, W // ws
    h = h.view(B, nWh, ws, nWw, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B * nWh * nWw, ws * ws, C)
    qkv = F.linear(h, sd[pre + ".attn.qkv.weight"], sd[pre + ".attn.qkv.bias"])
    qkv = qkv.view(-1, ws * ws, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    a = q @ k.transpose(-2, -1)
    bias = sd[pre + ".attn.relative_position_bias_table"][rel_pos_index(ws).reshape(-1)]
    a = a + bias.view(ws * ws, ws * ws, heads).permute(2, 0, 1).unsqueeze(0)
    if shift:
        m = shift_mask(H, W, ws, shift).to(a)
        a = (a.view(B, nWh * nWw, heads, ws * ws, ws * ws) + m[None, :, None]).view(-1, heads, ws * ws, ws * ws)
    a = torch.softmax(a, -1)
    o = (a @ v).transpose(1, 2).reshape(-1, ws * ws, C)
    o = F.linear(o, sd[pre + ".attn.proj.weight"], sd[pre + ".attn.proj.bias"])
    o = o.view(B, nWh, nWw, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    x = x + o
    h = F.layer_norm(x, (C,), sd[pre + ".norm2.weight"], sd[pre + ".norm2.bias"], 1e-5)
    h = F.linear(h, sd[pre + ".mlp.fc1.weight"], sd[pre + ".mlp.fc1.bias"])
    h = F.gelu(h)  # exact erf form, swinblock.py:14
    h = F.linear(h, sd[pre + ".mlp.fc2.weight"], sd[pre + ".mlp.fc2.bias"])
    return x + h


def _blocks(x, sd, pre, depth, heads, ws):
    for b in range(depth):
        x = swin_block(x, sd, f"{pre}.blocks.{b}", heads, ws, 0 if b % 2 == 0 else ws // 2)
    return x


def tower_encoder(x: Tensor, sd, pre: str, cfg: NetConfig):
    """Transformer_Encoder.forward, transformer.py:390-404.  x (B,C_g,H,W)."""
    B = x.shape[0]
    D, ws = cfg.enc_dim, cfg.window_size
    t = F.conv2d(x, sd[pre + ".patch_embed.proj.weight"], sd[pre + ".patch_embed.proj.bias"], stride=2)
    t = t.flatten(2).transpose(1, 2) + sd[pre + ".absolute_pos_embed"]          # :46, :394
    t = t.view(B, cfg.res0[0], cfg.res0[1], D)
    s0 = _blocks(t, sd, pre + ".layers.0", cfg.enc_depths[0], cfg.enc_heads[0], ws)
    # PatchMerging, transformer.py:76-96 (row-parity varies fastest in the concat order)
    m = torch.cat([s0[:, 0::2, 0::2], s0[:, 1::2, 0::2], s0[:, 0::2, 1::2], s0[:, 1::2, 1::2]], -1)
    m = F.layer_norm(m, (4 * D,), sd[pre + ".layers.1.downsample.norm.weight"],
                     sd[pre + ".layers.1.downsample.norm.bias"], 1e-6)
    m = F.linear(m, sd[pre + ".layers.1.downsample.reduction.weight"])
    s1 = _blocks(m, sd, pre + ".layers.1", cfg.enc_depths[1], cfg.enc_heads[1], ws)
    out = F.layer_norm(s1, (2 * D,), sd[pre + ".norm.weight"], sd[pre + ".norm.bias"], 1e-6)
    return out, [s0, s1]                                                        # skips are pre-norm, :399-402


def tower_decoder(x: Tensor, skips: List[Tensor], sd, pre: str, cfg: NetConfig) -> Tensor:
    """Transformer_Decoder.forward, transformer.py:466-474.  x (B,h1,w1,2D)."""
    D, ws = cfg.enc_dim, cfg.window_size
    x = F.linear(torch.cat([x, skips[1]], -1), sd[pre + ".concat_back_dim.0.weight"], sd[pre + ".concat_back_dim.0.bias"])
    x = _blocks(x, sd, pre + ".layers_up.0", cfg.enc_depths[1], cfg.enc_heads[1], ws)
    # PatchExpand, transformer.py:106-118: 'b h w (p1 p2 c) -> b (h p1) (w p2) c'
    x = F.linear(x, sd[pre + ".layers_up.0.upsample.expand.weight"])
    B, h, w, C4 = x.shape
    x = x.view(B, h, w, 2, 2, C4 // 4).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * h, 2 * w, C4 // 4)
    x = F.layer_norm(x, (D,), sd[pre + ".layers_up.0.upsample.norm.weight"], sd[pre + ".layers_up.0.upsample.norm.bias"], 1e-6)
    x = F.linear(torch.cat([x, skips[0]], -1), sd[pre + ".concat_back_dim.1.weight"], sd[pre + ".concat_back_dim.1.bias"])
    x = _blocks(x, sd, pre + ".layers_up.1", cfg.enc_depths[0], cfg.enc_heads[0], ws)
    return F.layer_norm(x, (D,), sd[pre + ".norm_up.weight"], sd[pre + ".norm_up.bias"], 1e-6)


def lgunet_forward(x: Tensor, sd: Dict[str, Tensor], cfg: NetConfig) -> Tensor:
    """LGUnet_all.forward, transformer.py:747-752.  x (B, sum C_in, H, W) -> (B, sum C_out, H, W)."""
    B = x.shape[0]
    G, D = cfg.groups, cfg.enc_dim
    parts = torch.split(x, list(cfg.inchans_list), 1)                           # Enc_net, :554-568
    feats, skips = [], []
    for g in range(G):
        f, s = tower_encoder(parts[g], sd, f"enc.enc_list.{g}", cfg)
        feats.append(f)
        skips.append(s)
    t = F.linear(torch.cat(feats, -1), sd["enc.proj.weight"], sd["enc.proj.bias"])
    h1, w1 = cfg.res1                                                           # LG_net, :698-712
    t = (t.reshape(B, h1 * w1, -1) + sd["net.pos_embed"]).view(B, h1, w1, -1)
    for l, depth in enumerate(cfg.lg_depths):
        t = _blocks(t, sd, f"net.layers.{l}", depth, cfg.lg_heads[l], cfg.window_size)
    t = F.linear(t, sd["dec.proj.weight"], sd["dec.proj.bias"])                 # Dec_net, :599-625
    chunks = torch.split(t, 2 * D, -1)
    means, stds = [], []
    for g in range(G):
        y = tower_decoder(chunks[g], skips[g], sd, f"dec.dec_list.{g}", cfg).permute(0, 3, 1, 2)
        y = F.conv_transpose2d(y, sd[f"dec.final_proj_list.{g}.weight"], sd[f"dec.final_proj_list.{g}.bias"], stride=2)
        half = y.shape[1] // 2
        means.append(y[:, :half])
        stds.append(y[:, half:])
    return torch.cat(means + stds, 1)


def to_torch(sd_np) -> Dict[str, Tensor]:
    return {k: torch.from_numpy(v) for k, v in sd_np.items()}
