"""Drop-in shells with the reference's module call surface, backed by the CUDA engine.

    LGUnet_all   networks_old/transformer.py:716-752   (constructor keywords :717-718, forward :747-752)
    LGUnet_all_1 networks/LGUnet_all.py:743-777        (the 0.25-degree forecast model; forward only, da_4dvar.py:555, 1329)
    VAE_lr       nf_model/vae.py:53-102                (encoder/decoder/decoder_hr/forward, .enc / .dec)

Parameters live in ordinary nn.Parameters under the reference's state_dict names (so reference checkpoints load
with load_state_dict and `strip "module."` logic of da_4dvar.py:592-601 keeps working); the arithmetic happens in
libvaevar.so.  `forward` is differentiable with respect to its INPUT through torch.autograd (da_4dvar.py:1244-1245);
weight gradients are deliberately not produced (the DA loop never reads them).
"""
from __future__ import annotations

import pathlib
from typing import Dict, Optional

import torch
import torch.nn as nn

from .config import DECODER_FULL, ENCODER_FULL, Net1Config, NetConfig
from .engine import Engine
from .forecast import ForecastNet
from .synth import make_state_dict, make_state_dict_net1


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's dotted parameter names."""


def _install(root: nn.Module, name: str, tensor: torch.Tensor):
    parts = name.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class _NetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, owner):
        eng = owner._engine()
        ctx.owner = owner
        ctx.save_for_backward(x)
        outs = [eng.net_forward(0, x[b].float()) for b in range(x.shape[0])]
        return torch.stack(outs, 0)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        eng = ctx.owner._engine()
        dx = [eng.net_vjp(0, x[b].float(), dy[b].float().contiguous()) for b in range(x.shape[0])]
        return torch.stack(dx, 0), None


class LGUnet_all(nn.Module):
    def __init__(self, rank=0, img_size=(32, 64), patch_size=(2, 2), stride=(2, 2), in_chans=20, out_chans=20,
                 enc_depths=(2, 2), enc_heads=(3, 6), lg_depths=(), lg_heads=(), inchans_list=(20,), outchans_list=(20,),
                 enc_dim=96, embed_dim=768, window_size=4, Weather_T=16, drop_rate=0., attn_drop_rate=0., drop_path=0.,
                 use_checkpoint=False, channel_num=37, inp_length=1, use_mlp=False, pre_norm=True, seed: int = 0):
        super().__init__()
        if rank:
            raise NotImplementedError("LoRA rank > 0 (swinblock.py:106-108) is a fine-tuning feature outside the DA path")
        if list(patch_size)[-2:] != [2, 2] or list(stride) != [2, 2] or inp_length != 1:
            raise NotImplementedError("only patch 2x2 / stride 2 / inp_length 1 (nf_model/parameters0_old.yaml) is built")
        if drop_rate or attn_drop_rate or drop_path:
            raise NotImplementedError("dropout / drop-path are training-only; the DA path runs in eval mode")
        ws = window_size if isinstance(window_size, int) else window_size[0]
        self.cfg = NetConfig(img_size=tuple(img_size), inchans_list=tuple(inchans_list), outchans_list=tuple(outchans_list),
                             enc_dim=enc_dim, embed_dim=embed_dim, window_size=ws, enc_depths=tuple(enc_depths),
                             enc_heads=tuple(enc_heads), lg_depths=tuple(lg_depths), lg_heads=tuple(lg_heads))
        for k, v in make_state_dict(self.cfg, seed=seed).items():
            _install(self, k, torch.from_numpy(v))
        self._eng: Optional[Engine] = None
        self._eng_version = None

    # any weight change invalidates the packed bf16 copy inside the engine
    def _version(self):
        return tuple(p._version for p in self.parameters()) + (str(next(self.parameters()).device),)

    def _engine(self) -> Engine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("vaevar_b200.LGUnet_all runs on a CUDA device only (no CPU fallback): call .to('cuda') first")
        v = self._version()
        if self._eng is None or v != self._eng_version:
            if self._eng is not None:
                self._eng.close()
            eng = Engine(self.cfg, None, T=1, use_graph=False, device=str(dev), dec_keep=0)
            eng.load_state_dict(0, {k: p.data for k, p in self.named_parameters()})
            eng.finalize()
            self._eng, self._eng_version = eng, v
        return self._eng

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = {k: v for k, v in state_dict.items() if "relative_position_index" not in k and "attn_mask" not in k}
        return super().load_state_dict(sd, strict=strict, **kw)

    def forward(self, data: torch.Tensor) -> torch.Tensor:
        """(B, sum C_in, H, W) -> (B, sum C_out, H, W); transformer.py:747-752."""
        return _NetFn.apply(data, self)


class LGUnet_all_1(nn.Module):
    """networks/LGUnet_all.py:743-777.  The DA driver builds it in init_model_forecast (da_4dvar.py:548-569), keeps it in eval
    mode under no_grad semantics (`integrate(..., detach=True)`, :666-681) and reads `model(x)[:, :69]`; the shell therefore
    produces no gradients at all.  `keep_out` (default: every output channel) lets a caller that only wants the mean channels skip
    the std half of the ConvTranspose2d head."""

    def __init__(self, img_size=(32, 64), patch_size=(1, 1, 1), stride=(2, 2), in_chans=20, out_chans=20, enc_depths=(2, 2),
                 enc_heads=(3, 6), lg_depths=(), lg_heads=(), inchans_list=(20,), outchans_list=(20,), enc_dim=96, embed_dim=768,
                 window_size=(4, 8), Weather_T=16, drop_rate=0., attn_drop_rate=0., drop_path=0., use_checkpoint=False, channel_num=37,
                 inp_length=1, use_mlp=False, pre_norm=True, seed: int = 0, keep_out: int = 0):
        super().__init__()
        if tuple(patch_size)[-2:] != (3, 2) or tuple(stride) != (2, 2) or inp_length != 1:
            raise NotImplementedError("only patch (3, 2) / stride (2, 2) / inp_length 1 (model_0.25degree/training_options.yaml:64-119) is built")
        if drop_rate or attn_drop_rate or drop_path:
            raise NotImplementedError("dropout / drop-path are training-only; the DA path runs in eval mode")
        if use_mlp or not pre_norm:
            raise NotImplementedError("use_mlp / post-norm blocks are not part of the shipped forecast model")
        self.cfg = Net1Config(img_size=tuple(img_size), patch_size=tuple(patch_size)[-2:], stride=tuple(stride),
                              inchans_list=tuple(inchans_list), outchans_list=tuple(outchans_list), enc_dim=enc_dim, embed_dim=embed_dim,
                              window_size=tuple(window_size), enc_depths=tuple(enc_depths), enc_heads=tuple(enc_heads),
                              lg_depths=tuple(lg_depths), lg_heads=tuple(lg_heads))
        self.keep_out = int(keep_out)
        for k, v in make_state_dict_net1(self.cfg, seed=seed).items():
            _install(self, k, torch.from_numpy(v))
        self._eng: Optional[ForecastNet] = None
        self._eng_version = None

    def _version(self):
        return tuple(p._version for p in self.parameters()) + (str(next(self.parameters()).device),)

    def _engine(self) -> ForecastNet:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("vaevar_b200.LGUnet_all_1 runs on a CUDA device only (no CPU fallback): call .to('cuda') first")
        v = self._version()
        if self._eng is None or v != self._eng_version:
            if self._eng is not None:
                self._eng.close()
            eng = ForecastNet(self.cfg, keep_out=self.keep_out, device=str(dev))
            eng.load_state_dict({k: p.data for k, p in self.named_parameters()})
            eng.finalize()
            self._eng, self._eng_version = eng, v
        return self._eng

    def forward(self, data: torch.Tensor, **kwargs) -> torch.Tensor:
        """(B, sum C_in, H, W) -> (B, keep_out or sum C_out, H, W); networks/LGUnet_all.py:772-777.  No autograd graph."""
        eng = self._engine()
        with torch.no_grad():
            return torch.stack([eng.forward(data[b].float()) for b in range(data.shape[0])], 0)

    def integrate(self, xa: torch.Tensor, steps: int = 1) -> torch.Tensor:
        """cyclic_4dvar.integrate(xa, self, steps) in one library call (da_4dvar.py:666-681): physical (69, H, W) in and out."""
        if self.keep_out != self.cfg.in_chans:
            raise RuntimeError("integrate needs keep_out == number of state channels (69): the model must map the state onto itself")
        return self._engine().integrate(xa, steps)


def _yaml_config(param_path: str) -> Dict[str, Dict]:
    """nf_model/<param_path>.yaml relative to the working directory, exactly as nf_model/vae.py:57 opens it;
    the shipped parameters0_old hyper-parameters are built in as a fallback so nothing needs /root/reference."""
    p = pathlib.Path("nf_model") / f"{param_path}.yaml"
    if p.exists():
        import yaml
        with open(p) as f:
            return yaml.load(f, Loader=yaml.FullLoader)
    if param_path != "parameters0_old":
        raise FileNotFoundError(str(p))
    return {"encoder": ENCODER_FULL.to_reference_kwargs(), "decoder": DECODER_FULL.to_reference_kwargs()}


class VAE_lr(nn.Module):
    """nf_model/vae.py:53-102."""

    def __init__(self, param_path: str = "parameters0_old", lora_rank: int = 0, build_encoder: bool = True):
        super().__init__()
        cfg = _yaml_config(param_path)
        self.param_encoder, self.param_decoder = dict(cfg["encoder"]), dict(cfg["decoder"])
        self.param_encoder["rank"] = lora_rank
        self.param_decoder["rank"] = lora_rank
        if build_encoder:
            self.enc = LGUnet_all(**self.param_encoder)
        self.dec = LGUnet_all(**self.param_decoder, seed=1)

    def encoder(self, x):
        return self.enc(x).chunk(2, dim=1)

    def sampling(self, mu, log_var):
        std = torch.exp(0.5 * log_var)
        return torch.randn_like(std).mul(std).add_(mu)

    def decoder(self, z):
        return self.dec(z)

    def decoder_hr(self, z):
        """nf_model/vae.py:87-90; the nearest up-sampling and its adjoint are the seam kernels (seams.py), not torch ops."""
        from .seams import interpolate_nearest
        return interpolate_nearest(self.dec(z), (721, 1440))

    def forward(self, x):
        mu, log_var = self.encoder(x)
        return self.decoder(self.sampling(mu, log_var)), mu, log_var
