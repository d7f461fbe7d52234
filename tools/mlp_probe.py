"""Run the fused tower MLP kernels (mlp_fused.cuh) at one shape a few times: event timing, and the target of ncu captures.
    python tools/mlp_probe.py --shape 8192,96,6 [--reps 20] [--which fwd,bwd]"""
import argparse
import ctypes as C
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch

from vaevar_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="8192,96,6")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--which", default="fwd,bwd")
a = ap.parse_args()
rows, D, B = map(int, a.shape.split(","))
lib = _lib.load()
dev = "cuda:0"
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
dt = torch.float16
x1 = torch.randn(B, rows, D, device=dev); W1 = (torch.randn(B, 4 * D, D, device=dev) * 0.08).to(dt); W2 = (torch.randn(B, D, 4 * D, device=dev) * 0.05).to(dt)
b1 = torch.randn(B, 4 * D, device=dev); b2 = torch.randn(B, D, device=dev); shift = x1.mean(-1).contiguous()
u = torch.empty(B, rows, 4 * D, device=dev, dtype=dt); out = torch.empty(B, rows, D, device=dev); o16 = torch.empty(B, rows, D, device=dev, dtype=dt)
stats = torch.empty(B, rows, 2, device=dev)
dyb = torch.randn(B, rows, D, device=dev).bfloat16(); dres = torch.randn(B, rows, D, device=dev); gamma = torch.ones(B, D, device=dev)
W2T = W2.float().transpose(1, 2).contiguous().bfloat16(); W1T = W1.float().transpose(1, 2).contiguous().bfloat16()
dx = torch.empty(B, rows, D, device=dev); dxb = torch.empty(B, rows, D, device=dev, dtype=torch.bfloat16)
fns = {"fwd": lambda: lib.vv_test_mlp_fwd(P(x1), P(W1), P(W2), P(b1), P(b2), rows, B, D, 1, 1e-5, P(u), P(out), P(o16), P(shift), P(stats), st),
       "bwd": lambda: lib.vv_test_mlp_bwd(P(dyb), P(u), P(x1), P(W2T), P(W1T), P(gamma), P(dres), rows, B, D, 1, 1e-5, P(dx), P(dxb), st)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in a.which.split(","):
    fn = fns[name]
    for _ in range(3):
        _lib.check(fn())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / a.reps
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"mlp {name} {rows}x{D}x{B}: back to back {warm*1e3:.1f} us | L2 flushed median {ts[len(ts)//2]*1e3:.1f} us (min {ts[0]*1e3:.1f}) "
          f"= {16.0*rows*D*D*B/ts[len(ts)//2]/1e9:.0f} TFLOP/s", flush=True)
