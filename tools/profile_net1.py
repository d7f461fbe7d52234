"""Per-launch steady-state profile of one LGUnet_all_1 application at the shipped 0.25-degree size (69 x 721 x 1440), grouped by
kernel and shape, with the algorithmic TFLOP/s of the GEMM and attention launches.
    python tools/profile_net1.py [--mid] [--json out.json]"""
import argparse
import collections
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch

from vaevar_b200.config import FORECAST_FULL, FORECAST_MID
from vaevar_b200.forecast import ForecastNet
from vaevar_b200.synth import make_state_dict_net1

ap = argparse.ArgumentParser()
ap.add_argument("--mid", action="store_true")
ap.add_argument("--json", default="")
a = ap.parse_args()
cfg = FORECAST_MID if a.mid else FORECAST_FULL
net = ForecastNet(cfg, keep_out=69)
net.load_state_dict(make_state_dict_net1(cfg, seed=1))
net.finalize()
x = torch.randn(69, *cfg.img_size, generator=torch.Generator().manual_seed(3)).cuda()
for _ in range(2):
    y = net.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    y = net.forward(x)
e1.record(); torch.cuda.synchronize()
ms_app = e0.elapsed_time(e1) / 3
ops = net.profile_ops(3)
agg = collections.OrderedDict()
for o in ops:
    k = (o["kind"],) + tuple(o["shape"])
    r = agg.setdefault(k, [0, 0.0, 0.0])
    r[0] += 1; r[1] += o["ms"]; r[2] += o["flop"]
tot = sum(o["ms"] for o in ops); fl = sum(o["flop"] for o in ops)
print(f"LGUnet_all_1 {cfg.img_size}: {ms_app:.2f} ms per application ({net.last_launch_count} launches, {net.device_bytes / 2**30:.2f} GiB); "
      f"sum of per-op times {tot:.2f} ms; GEMM + attention flops {fl / 1e12:.2f} TFLOP -> {fl / 1e9 / ms_app:.0f} TFLOP/s")
rows = []
for k, (n, ms, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    rows.append({"kernel": k[0], "shape": list(k[1:]), "launches": n, "ms": round(ms, 3), "share": round(ms / tot, 4), "TFLOP/s": round(f / 1e9 / ms, 1) if f else None})
    print(f"  {ms:8.3f} ms {100 * ms / tot:5.1f}%  n={n:3d}  {k[0]:14s} {str(list(k[1:])):32s} " + (f"{f / 1e9 / ms:7.0f} TFLOP/s" if f else ""))
if a.json:
    pathlib.Path(a.json).write_text(json.dumps({"img_size": list(cfg.img_size), "ms_per_application": ms_app, "launches": net.last_launch_count,
                                                "device_GiB": net.device_bytes / 2**30, "tflop": fl / 1e12, "rows": rows}))
