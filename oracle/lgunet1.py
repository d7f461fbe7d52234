"""ORACLE (test infrastructure, never on the product path): CPU fp32 restatement of the reference's forecast network
`LGUnet_all_1` (networks/LGUnet_all.py:743-777) as a pure function of a state_dict with the reference's parameter names.
Forward only: the DA cycle never differentiates it (`integrate(xa, forecast_model, 1)` with detach, da_4dvar.py:1329, 666-681).

    LGUnet_all_1.forward                 networks/LGUnet_all.py:772-777
    Enc_net / Transformer_Encoder        :553-590, :345-412   (PatchEmbed :14-50, PatchMerging :64-98)
    LG_net                               :653-739             (first stage = one window over the whole grid, then shifted 6x12 windows)
    Dec_net / Transformer_Decoder        :592-650, :414-477   (PatchExpand :101-118, ConvTranspose2d head, mean / std halves)
    Windowattn_block                     networks/utils/Blocks.py:103-159   (pre-norm)
    SD_attn (dilation 1)                 networks/utils/Attention.py:467-664
    rope2                                networks/utils/positional_encodings.py:230-268
    window_partition / window_reverse    networks/utils/utils.py:82-135

"parity unpinned" by the reference's own tests (it has none); pinned against the reference module itself run on CPU in the build
container (tools/make_golden.py::golden_lgunet1 -> tests/golden/net1_small.npz).

Quirks kept: every LayerNorm has eps 1e-6; the shift mask is 0 / -inf and, because the third longitude slice of create_mask
(`slice(0, None)`, Attention.py:528-530) overwrites the first two, depends on latitude only; the forward roll is keyed on the longitude
shift, the backward roll on the latitude shift (:560, :641); a block whose window spans the whole width gets no mask even when shifted
(:553); q is rotated, then scaled (:600-606); the ConvTranspose2d head has kernel (3, 2) / stride 2, so adjacent patch rows overlap-add.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass(frozen=True)
class Net1Config:
    """Constructor arguments of LGUnet_all_1 that shape the computation (output/model/model_0.25degree/training_options.yaml:64-119)."""
    img_size: Tuple[int, int] = (721, 1440)
    patch_size: Tuple[int, int] = (3, 2)
    stride: Tuple[int, int] = (2, 2)
    inchans_list: Tuple[int, ...] = (4, 13, 13, 13, 13, 13)
    outchans_list: Tuple[int, ...] = (8, 26, 26, 26, 26, 26)
    enc_dim: int = 96
    embed_dim: int = 1152
    window_size: Tuple[int, int] = (6, 12)
    enc_depths: Tuple[int, ...] = (2, 2, 2)
    enc_heads: Tuple[int, ...] = (3, 6, 6)
    lg_depths: Tuple[int, ...] = (4, 4, 4)
    lg_heads: Tuple[int, ...] = (6, 6, 6)

    @property
    def patches(self) -> Tuple[int, int]:
        return ((self.img_size[0] - self.patch_size[0]) // self.stride[0] + 1, (self.img_size[1] - self.patch_size[1]) // self.stride[1] + 1)

    def to_reference_kwargs(self) -> Dict:
        return dict(img_size=list(self.img_size), patch_size=list(self.patch_size), stride=list(self.stride),
                    inchans_list=list(self.inchans_list), outchans_list=list(self.outchans_list), in_chans=sum(self.inchans_list),
                    out_chans=sum(self.outchans_list), enc_dim=self.enc_dim, embed_dim=self.embed_dim, window_size=list(self.window_size),
                    enc_depths=list(self.enc_depths), enc_heads=list(self.enc_heads), lg_depths=list(self.lg_depths),
                    lg_heads=list(self.lg_heads), Weather_T=1, drop_path=0.0, use_checkpoint=False, inp_length=1, use_mlp=False)


EPS = 1e-6


def _ln(x: Tensor, sd, prefix: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], EPS)


def _lin(x: Tensor, sd, prefix: str) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def rope2_tables(shape: Sequence[int], dim: int):
    """positional_encodings.py:231-252: (sin1, cos1, sin2, cos2), each (Mh, Mw, dim // 4) for dim % 4 == 0."""
    c0, c1 = torch.arange(shape[0]), torch.arange(shape[1])
    coords = torch.stack(torch.meshgrid([c0, c1], indexing="ij")).reshape(2, -1)
    half = dim // 2
    d1, d2 = half // 2, half - half // 2
    inv1 = 10000 ** -(torch.arange(0, d1) / d1)
    inv2 = 10000 ** -(torch.arange(0, d2) / d2)
    s1, s2 = coords[0].unsqueeze(-1) * inv1, coords[1].unsqueeze(-1) * inv2
    return (torch.sin(s1).reshape(*shape, d1), torch.cos(s1).reshape(*shape, d1),
            torch.sin(s2).reshape(*shape, d2), torch.cos(s2).reshape(*shape, d2), d1, d2)


def rope2(x: Tensor, tab) -> Tensor:
    """positional_encodings.py:255-268 on x (..., Mh, Mw, dim): rows rotate the (x11, x12) pair, columns the (x21, x22) pair."""
    sin1, cos1, sin2, cos2, d1, d2 = tab
    x11, x21, x12, x22 = x.split([d1, d2, d1, d2], dim=-1)
    return torch.cat([x11 * cos1 - x12 * sin1, x21 * cos2 - x22 * sin2, x12 * cos1 + x11 * sin1, x22 * cos2 + x21 * sin2], dim=-1)


def _partition(x: Tensor, win) -> Tensor:
    B, H, W, C = x.shape
    x = x.view(B, H // win[0], win[0], W // win[1], win[1], C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, win[0], win[1], C)


def _reverse(w: Tensor, win, H: int, W: int) -> Tensor:
    B = w.shape[0] // ((H // win[0]) * (W // win[1]))
    x = w.view(B, H // win[0], W // win[1], win[0], win[1], -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def shift_mask(H: int, W: int, win, shift) -> Tensor:
    """SD_attn.create_mask for 2-D windows (Attention.py:520-548): (nW, N, N) of 0 / -inf; latitude bands only."""
    img = torch.zeros(1, H, W, 1)
    cnt = 0
    for hs in (slice(0, -win[0]), slice(-win[0], -shift[0]), slice(-shift[0], None)):
        for ws in (slice(0, -win[1]), slice(-win[1], 0), slice(0, None)):
            img[:, hs, ws, :] = cnt
            cnt += 1
    m = _partition(img, win).reshape(-1, win[0] * win[1])
    d = m.unsqueeze(1) - m.unsqueeze(2)
    return d.masked_fill(d != 0, -torch.inf).masked_fill(d == 0, 0.0)


def sd_attn(x: Tensor, sd, prefix: str, heads: int, win, shift) -> Tensor:
    """SD_attn.forward (Attention.py:551-650) on x (B, H, W, C) with dilation 1."""
    B0, H, W, C = x.shape
    hd = C // heads
    mask = None if (shift[1] == 0 or win[1] == W) else shift_mask(H, W, win, shift)
    xs = torch.roll(x, shifts=(-shift[0], -shift[1]), dims=(1, 2)) if shift[1] > 0 else x
    xw = _partition(xs, win).reshape(-1, win[0] * win[1], C)
    B_, N, _ = xw.shape
    qkv = _lin(xw, sd, prefix + ".qkv").reshape(B_, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    tab = rope2_tables(win, hd)
    q = rope2(q.reshape(-1, win[0], win[1], hd), tab).reshape(B_, heads, N, hd)
    k = rope2(k.reshape(-1, win[0], win[1], hd), tab).reshape(B_, heads, N, hd)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    xs = _reverse(out.reshape(-1, win[0], win[1], C), win, H, W)
    if shift[0] > 0:
        xs = torch.roll(xs, shifts=(shift[0], shift[1]), dims=(1, 2))
    return _lin(xs, sd, prefix + ".proj")


def block(x: Tensor, sd, prefix: str, heads: int, win, shift) -> Tensor:
    """Windowattn_block.forward, pre-norm (Blocks.py:142-157)."""
    x = x + sd_attn(_ln(x, sd, prefix + ".norm"), sd, prefix + ".attn", heads, win, shift)
    h = F.gelu(_lin(_ln(x, sd, prefix + ".norm2"), sd, prefix + ".mlp.fc1"))
    return x + _lin(h, sd, prefix + ".mlp.fc2")


def _stage(x: Tensor, sd, prefix: str, depth: int, heads: int, win, shifted: bool = True) -> Tensor:
    for i in range(depth):
        shift = (win[0] // 2, win[1] // 2) if (shifted and i % 2 == 1) else (0, 0)
        x = block(x, sd, f"{prefix}.blocks.{i}", heads, win, shift)
    return x


def patch_merging(x: Tensor, sd, prefix: str) -> Tensor:
    """LGUnet_all.py:80-98."""
    B, H, W, C = x.shape
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    return F.linear(_ln(x, sd, prefix + ".norm"), sd[prefix + ".reduction.weight"])


def patch_expand(x: Tensor, sd, prefix: str) -> Tensor:
    """LGUnet_all.py:107-118: Linear(dim -> 2 dim), pixel shuffle 'b h w (p1 p2 c) -> b (h p1) (w p2) c', LayerNorm(dim / 2)."""
    x = F.linear(x, sd[prefix + ".expand.weight"])
    B, H, W, C = x.shape
    x = x.view(B, H, W, 2, 2, C // 4).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, C // 4)
    return _ln(x, sd, prefix + ".norm")


def encoder_tower(x: Tensor, sd, prefix: str, cfg: Net1Config) -> Tuple[Tensor, List[Tensor]]:
    """Transformer_Encoder.forward (LGUnet_all.py:396-412)."""
    B = x.shape[0]
    H0, W0 = cfg.patches
    x = F.conv2d(x, sd[prefix + ".patch_embed.proj.weight"], sd[prefix + ".patch_embed.proj.bias"], stride=cfg.stride)
    x = x.flatten(2).transpose(1, 2) + sd[prefix + ".absolute_pos_embed"]
    x = x.view(B, H0, W0, -1)
    skips = []
    for i, depth in enumerate(cfg.enc_depths):
        if i > 0:
            x = patch_merging(x, sd, f"{prefix}.layers.{i}.downsample")
        x = _stage(x, sd, f"{prefix}.layers.{i}", depth, cfg.enc_heads[i], cfg.window_size)
        skips.append(x)
    return _ln(x, sd, prefix + ".norm"), skips


def decoder_tower(x: Tensor, skips: List[Tensor], sd, prefix: str, cfg: Net1Config) -> Tensor:
    """Transformer_Decoder.forward (LGUnet_all.py:466-475)."""
    n = len(cfg.enc_depths)
    for inx in range(n):
        lvl = n - 1 - inx
        x = _lin(torch.cat([x, skips[lvl]], -1), sd, f"{prefix}.concat_back_dim.{inx}")
        x = _stage(x, sd, f"{prefix}.layers_up.{inx}", cfg.enc_depths[lvl], cfg.enc_heads[lvl], cfg.window_size)
        if inx < n - 1:
            x = patch_expand(x, sd, f"{prefix}.layers_up.{inx}.upsample")
    return _ln(x, sd, prefix + ".norm_up")


def lgunet1_forward(x: Tensor, sd: Dict[str, Tensor], cfg: Net1Config) -> Tensor:
    """(B, sum C_in, H, W) -> (B, sum C_out, H, W); LGUnet_all_1.forward."""
    G = len(cfg.inchans_list)
    lasts, skips = [], []
    for g, xg in enumerate(torch.split(x, list(cfg.inchans_list), dim=1)):                     # Enc_net.forward, :575-590
        last, sk = encoder_tower(xg, sd, f"enc.enc_list.{g}", cfg)
        lasts.append(last); skips.append(sk)
    h = _lin(torch.cat(lasts, dim=-1), sd, "enc.proj")
    B, H, W, C = h.shape                                                                       # LG_net.forward, :722-739
    h = (h.reshape(B, -1, C) + sd["net.pos_embed"]).view(B, H, W, C)
    for i, depth in enumerate(cfg.lg_depths):
        if i == 0:
            h = _stage(h, sd, "net.layers.0", depth, cfg.lg_heads[0], (H, W), shifted=False)   # one window = the whole grid
        else:
            h = _stage(h, sd, f"net.layers.{i}", depth, cfg.lg_heads[i], cfg.window_size)
    top = cfg.enc_dim * 2 ** (len(cfg.enc_depths) - 1)
    means, stds = [], []
    for g, hg in enumerate(torch.split(_lin(h, sd, "dec.proj"), top, dim=-1)):                 # Dec_net.forward, :625-650
        t = decoder_tower(hg, skips[g], sd, f"dec.dec_list.{g}", cfg).permute(0, 3, 1, 2)
        o = F.conv_transpose2d(t, sd[f"dec.final_proj_list.{g}.weight"], sd[f"dec.final_proj_list.{g}.bias"], stride=cfg.stride)
        means.append(o[:, : o.shape[1] // 2]); stds.append(o[:, o.shape[1] // 2:])
    return torch.cat(means + stds, dim=1)


NET1_SMALL = Net1Config(img_size=(49, 96), enc_dim=16, embed_dim=64, enc_heads=(2, 2, 2), lg_depths=(2, 2), lg_heads=(2, 2))


def synth_state_dict(shapes: Dict[str, Sequence[int]], seed: int = 0) -> Dict[str, Tensor]:
    """Deterministic weights by parameter name (the fixture stores names and shapes only): LayerNorm weights around 1, other weights
    N(0, 0.08), every bias / embedding N(0, 0.1), so no parameter is silent."""
    import hashlib

    import numpy as np
    out = {}
    for name in sorted(shapes):
        shape = tuple(int(v) for v in shapes[name])
        h = int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:8], "little")
        r = np.random.Generator(np.random.PCG64(h)).standard_normal(shape, dtype=np.float32)
        if name.endswith("weight") and len(shape) == 1:
            v = 1.0 + 0.1 * r
        elif name.endswith("weight"):
            v = 0.08 * r
        else:
            v = 0.1 * r
        out[name] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return out
