"""Hyper-parameters of one U-shaped Swin network application (`LGUnet_all`).

Mirrors the keyword surface of networks_old/transformer.py:717-718 and the two
blocks of nf_model/parameters0_old.yaml (decoder: :49-96, encoder/flow: :1-48).
Only the knobs the shipped configs exercise are kept: patch 2x2 / stride 2,
two tower stages (enc_depths [2,2]), window 4, mlp_ratio 4, no dropout.
"""
from __future__ import annotations

import dataclasses
import json
import pathlib
from typing import Dict, List, Tuple

import numpy as np

_DATA = pathlib.Path(__file__).resolve().parent / "data"


@dataclasses.dataclass(frozen=True)
class NetConfig:
    img_size: Tuple[int, int] = (128, 256)
    inchans_list: Tuple[int, ...] = (2, 6, 6, 6, 6, 6)
    outchans_list: Tuple[int, ...] = (4, 13, 13, 13, 13, 13)
    enc_dim: int = 96
    embed_dim: int = 1152
    window_size: int = 4
    enc_depths: Tuple[int, int] = (2, 2)
    enc_heads: Tuple[int, int] = (3, 6)
    lg_depths: Tuple[int, ...] = (4, 4, 4)
    lg_heads: Tuple[int, ...] = (6, 6, 6)

    # ---- derived ---------------------------------------------------------
    @property
    def groups(self) -> int:
        return len(self.inchans_list)

    @property
    def in_chans(self) -> int:
        return int(sum(self.inchans_list))

    @property
    def out_chans(self) -> int:
        return int(sum(self.outchans_list))

    @property
    def res0(self) -> Tuple[int, int]:  # tower stage-0 token grid
        return (self.img_size[0] // 2, self.img_size[1] // 2)

    @property
    def res1(self) -> Tuple[int, int]:  # tower stage-1 / trunk token grid
        return (self.img_size[0] // 4, self.img_size[1] // 4)

    @property
    def mean_half(self) -> int:
        """Channels in the leading ("mean") half of the output, transformer.py:616-623."""
        return int(sum(c // 2 for c in self.outchans_list))

    def output_channel_map(self) -> List[Tuple[int, int]]:
        """For every output channel (reference order) the (group, channel-in-group)
        it comes from: all first halves, then all second halves (transformer.py:616-623)."""
        first, second = [], []
        for g, c in enumerate(self.outchans_list):
            h = c // 2
            first += [(g, k) for k in range(h)]
            second += [(g, k) for k in range(h, c)]
        return first + second

    def to_reference_kwargs(self) -> Dict:
        """Keyword dict accepted by the reference constructor (transformer.py:717-718)."""
        return dict(
            img_size=list(self.img_size), patch_size=[2, 2], stride=[2, 2],
            inchans_list=list(self.inchans_list), outchans_list=list(self.outchans_list),
            in_chans=138, out_chans=138, enc_dim=self.enc_dim, embed_dim=self.embed_dim,
            window_size=self.window_size, enc_depths=list(self.enc_depths),
            enc_heads=list(self.enc_heads), lg_depths=list(self.lg_depths),
            lg_heads=list(self.lg_heads), Weather_T=1, drop_path=0.0,
            use_checkpoint=False, inp_length=1, use_mlp=False)

    @staticmethod
    def from_reference_kwargs(kw: Dict) -> "NetConfig":
        return NetConfig(
            img_size=tuple(kw["img_size"]), inchans_list=tuple(kw["inchans_list"]),
            outchans_list=tuple(kw["outchans_list"]), enc_dim=kw["enc_dim"],
            embed_dim=kw["embed_dim"], window_size=kw["window_size"],
            enc_depths=tuple(kw["enc_depths"]), enc_heads=tuple(kw["enc_heads"]),
            lg_depths=tuple(kw["lg_depths"]), lg_heads=tuple(kw["lg_heads"]))


# nf_model/parameters0_old.yaml:49-96 -- the VAE decoder D (32 latent -> 69 state channels)
DECODER_FULL = NetConfig()
# nf_model/parameters0_old.yaml:1-48 with outchans_list [8,26x5] -- the forecast operator M
# (69 in, 138 out, first 69 kept, da_4dvar.py:674)
FLOW_FULL = NetConfig(inchans_list=(4, 13, 13, 13, 13, 13), outchans_list=(8, 26, 26, 26, 26, 26))
# nf_model/parameters0_old.yaml:1-48 -- the VAE encoder (69 -> mu 32 ++ log-var 32)
ENCODER_FULL = NetConfig(inchans_list=(4, 13, 13, 13, 13, 13), outchans_list=(4, 12, 12, 12, 12, 12))


def small(cfg: NetConfig, img=(32, 64), enc_dim=64, embed_dim=384, enc_heads=(2, 4),
          lg_depths=(2, 2), lg_heads=(2, 2)) -> NetConfig:
    """Shrunken twin of a full config (same channel lists, same head dims 32 / 192)
    that the CPU oracle finishes in a fraction of a second."""
    return dataclasses.replace(cfg, img_size=img, enc_dim=enc_dim, embed_dim=embed_dim,
                               enc_heads=enc_heads, lg_depths=lg_depths, lg_heads=lg_heads)


@dataclasses.dataclass(frozen=True)
class Net1Config:
    """Constructor arguments of the forecast network `LGUnet_all_1` that shape the computation (networks/LGUnet_all.py:743-744;
    the shipped values: output/model/model_0.25degree/training_options.yaml:64-119)."""
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
    def groups(self) -> int:
        return len(self.inchans_list)

    @property
    def in_chans(self) -> int:
        return int(sum(self.inchans_list))

    @property
    def out_chans(self) -> int:
        return int(sum(self.outchans_list))

    @property
    def patches(self) -> Tuple[int, int]:
        """Token grid of tower level 0 (PatchEmbed, networks/LGUnet_all.py:14-50)."""
        return ((self.img_size[0] - self.patch_size[0]) // self.stride[0] + 1, (self.img_size[1] - self.patch_size[1]) // self.stride[1] + 1)

    def level_grid(self, level: int) -> Tuple[int, int]:
        h, w = self.patches
        return (h >> level, w >> level)

    def to_reference_kwargs(self) -> Dict:
        return dict(img_size=list(self.img_size), patch_size=list(self.patch_size), stride=list(self.stride),
                    inchans_list=list(self.inchans_list), outchans_list=list(self.outchans_list), in_chans=self.in_chans,
                    out_chans=self.out_chans, enc_dim=self.enc_dim, embed_dim=self.embed_dim, window_size=list(self.window_size),
                    enc_depths=list(self.enc_depths), enc_heads=list(self.enc_heads), lg_depths=list(self.lg_depths),
                    lg_heads=list(self.lg_heads), Weather_T=1, drop_path=0.0, use_checkpoint=False, inp_length=1, use_mlp=False)


# output/model/model_0.25degree/training_options.yaml:64-119 -- the 0.25-degree forecast model of the DA cycle (da_4dvar.py:555, 1329)
FORECAST_FULL = Net1Config()
# a shrunken twin with the real geometry (patch (3, 2) / stride 2, 6 x 12 windows, three tower levels, head widths 32 / 32 / 64 / 192)
# that the reference module finishes on CPU in seconds: 97 x 192 image -> 48 x 96, 24 x 48, 12 x 24 tokens
FORECAST_MID = Net1Config(img_size=(97, 192), embed_dim=384, lg_depths=(2, 2), lg_heads=(2, 2))


def era5_stats():
    """(mean[69], std[69], stdTr[69]) as float64 numpy arrays.
    Values: da_4dvar.py:641-643 and :1181 (extracted by tools/extract_constants.py)."""
    d = json.loads((_DATA / "era5_stats.json").read_text())
    return (np.asarray(d["mean"], np.float64), np.asarray(d["std"], np.float64),
            np.asarray(d["stdTr"], np.float64))
