"""Device times of the native-resolution seam kernels at the reference's geometry (69 x 721 x 1440 <-> 69 x 128 x 256), CUDA events,
L2 flushed between launches, against the measured HBM copy bandwidth.
    python tools/seam_bench.py [--reps 20] > profiles/r1_seams.txt"""
import argparse
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from vaevar_b200 import seams
from vaevar_b200.engine import compact_mask

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
peak = 6550.0
pk = ROOT / "MEASURED_PEAKS.json"
if pk.exists():
    d = json.loads(pk.read_text())
    peak = float(d.get("hbm_gbs", peak))
dev = "cuda:0"
C, lo, hi = 69, (128, 256), (721, 1440)
nlo, nhi = C * lo[0] * lo[1], C * hi[0] * hi[1]
torch.manual_seed(0)
x_lo, x_hi = torch.randn(C, *lo, device=dev), torch.randn(C, *hi, device=dev)
mean, std = torch.randn(C, device=dev), torch.rand(C, device=dev) + 0.5
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


H = torch.zeros(hi[0] * hi[1], device=dev)
H[torch.randperm(hi[0] * hi[1], device=dev)[: int(0.1 * hi[0] * hi[1])]] = 1.0
H = H.reshape(1, *hi).expand(C, *hi).contiguous()
idx, y, rinv = compact_mask(H, x_hi + 0.1, torch.full_like(x_hi, 0.5))
n_obs = idx.numel()
rows = [
    ("up 128x256 -> 721x1440, * std + mean (da_4dvar.py:679,681)", lambda: seams.resample_nearest(x_lo, hi, 2, mean, std), 4 * (nlo + nhi)),
    ("down 721x1440 -> 128x256, (x - mean) / std (:667,671)", lambda: seams.resample_nearest(x_hi, lo, 1, mean, std), 4 * (nlo + nlo)),
    ("adjoint of up (sum of ~31.7 cotangents per cell)", lambda: seams.resample_nearest_adjoint(x_hi, lo, 2, std), 4 * (nhi + nlo)),
    ("adjoint of down (scatter into the 721x1440 grid)", lambda: seams.resample_nearest_adjoint(x_lo, hi, 1, std), 4 * (nlo + nhi)),
    (f"obs term, J only, {n_obs} obs (10 % of columns, :1207)", lambda: seams.obs_term(x_hi, idx, y, rinv, 1.0, False), 16 * n_obs),
    (f"obs term, J + gradient field (zero-fill {4 * nhi / 1e6:.0f} MB + scatter)", lambda: seams.obs_term(x_hi, idx, y, rinv, 1.0, True),
     20 * n_obs + 4 * nhi),
]
print(f"# seam kernels, C={C}; algorithmic bytes / median device time over {a.reps} launches (L2 flushed); HBM peak {peak:.0f} GB/s")
for name, fn, nbytes in rows:
    ms = timeit(fn)
    print(f"{name:75s} {ms * 1e3:8.1f} us  {nbytes / 1e6:8.1f} MB  {nbytes / ms / 1e6:7.0f} GB/s  frac {nbytes / ms / 1e6 / peak:.2f}")
