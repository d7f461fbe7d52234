"""Generate tests/golden/seams.npz with torch's own F.interpolate / autograd on CPU - the arithmetic the reference calls at
nf_model/vae.py:90 and da_4dvar.py:667-681.  Runs in the build container; the fixture is committed.
    python tools/make_golden_seams.py
"""
import pathlib

import numpy as np
import torch
import torch.nn.functional as F

ROOT = pathlib.Path(__file__).resolve().parents[1]
torch.manual_seed(0)
out = {}
# index tables at the reference's size pairs: F.interpolate of an index image
for name, (hi, wi, ho, wo) in {"up": (128, 256, 721, 1440), "down": (721, 1440, 128, 256)}.items():
    img = torch.arange(hi * wi, dtype=torch.float32).reshape(1, 1, hi, wi)          # < 2^24: exact in float32
    y = F.interpolate(img, (ho, wo))[0, 0].to(torch.int64)
    out[f"{name}_rows"] = (y[:, 0] // wi).numpy().astype(np.int32)
    out[f"{name}_cols"] = (y[0, :] % wi).numpy().astype(np.int32)
    assert torch.equal(y, (y[:, :1] // wi) * wi + (y[:1, :] % wi))                  # separable
# integrate()'s two seams with their (de)normalisation, forward and backward, on a small grid with the same non-integer ratios
C, lo, hi = 3, (8, 16), (45, 90)
mean = torch.tensor([1.5, -20.0, 300.0]); std = torch.tensor([2.0, 7.5, 0.125 * 3])
xa = (torch.randn(C, *hi) * std.reshape(-1, 1, 1) + mean.reshape(-1, 1, 1)).requires_grad_(True)
za = (xa - mean.reshape(-1, 1, 1)) / std.reshape(-1, 1, 1)                         # da_4dvar.py:667
z = F.interpolate(za.unsqueeze(0), lo)                                              # :670-671
g_lo = torch.randn_like(z)
z.backward(g_lo)
out.update(C=np.int32(C), lo=np.array(lo, np.int32), hi=np.array(hi, np.int32), mean=mean.numpy(), std=std.numpy(),
           xa=xa.detach().numpy(), down_norm=z[0].detach().numpy(), g_lo=g_lo[0].numpy(), down_norm_grad=xa.grad.numpy())
zl = torch.randn(1, C, *lo, requires_grad=True)
zu = F.interpolate(zl, hi)                                                          # :678-679
xp = zu.reshape(C, *hi) * std.reshape(-1, 1, 1) + mean.reshape(-1, 1, 1)            # :681
g_hi = torch.randn_like(xp)
xp.backward(g_hi)
out.update(zl=zl[0].detach().numpy(), up_denorm=xp.detach().numpy(), g_hi=g_hi.numpy(), up_denorm_grad=zl.grad[0].numpy())
# plain up-sampling (decoder_hr, vae.py:90) incl. the exact-doubling and identity branches
for tag, size in {"plain": hi, "double": (16, 32), "same": lo}.items():
    a = torch.randn(1, C, *lo, requires_grad=True)
    b = F.interpolate(a, size)
    g = torch.randn_like(b)
    b.backward(g)
    out.update({f"{tag}_in": a[0].detach().numpy(), f"{tag}_out": b[0].detach().numpy(), f"{tag}_g": g[0].numpy(),
                f"{tag}_grad": a.grad[0].numpy()})
# observation term on the analysis grid (da_4dvar.py:1207), float32 like the reference, and float64 for the tolerance
T = 2
x = (torch.randn(T, C, *hi)).requires_grad_(True)
yo = torch.randn(T, C, *hi)
Hm = torch.zeros(hi[0] * hi[1]); Hm[torch.randperm(hi[0] * hi[1])[:400]] = 1.0
H = Hm.reshape(1, 1, *hi).expand(T, C, *hi).contiguous()
R = (0.05 + torch.rand(T, C, 1, 1)).expand(T, C, *hi).contiguous()
loss = torch.sum(H * (x - yo) ** 2 / R) / 2
loss.backward()
out.update(obs_x=x.detach().numpy(), obs_yo=yo.numpy(), obs_H=H.numpy(), obs_R=R.numpy(), obs_J=np.float64(loss.item()),
           obs_grad=x.grad.numpy(),
           obs_J64=np.float64((H.double() * (x.detach().double() - yo.double()) ** 2 / R.double()).sum().item() / 2))
np.savez_compressed(ROOT / "tests" / "golden" / "seams.npz", **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
