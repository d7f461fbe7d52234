"""Run the fused norm1 + qkv kernel (mlp_fused.cuh, MLP_LIN mode) at one shape a few times: event timing, and the target of ncu captures.
    python tools/lin_probe.py --shape 8192,96,6 [--reps 20]"""
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
a = ap.parse_args()
rows, D, B = map(int, a.shape.split(","))
lib = _lib.load()
dev = "cuda:0"
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
x = torch.randn(B, rows, D, device=dev); W = (torch.randn(B, 3 * D, D, device=dev) * 0.08).half(); bias = torch.randn(B, 3 * D, device=dev)
out = torch.empty(B, rows, 3 * D, device=dev, dtype=torch.float16)
fn = lambda: lib.vv_test_lin_fwd(P(x), P(W), P(bias), rows, B, D, 3 * D, 1, 1e-5, P(out), st)
for _ in range(3):
    _lib.check(fn())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    fn()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / a.reps * 1e3
nbytes = B * rows * (D * 4 + 3 * D * 2)
print(f"norm1+qkv {rows}x{D}x{B}: back to back {us:.1f} us = {2.0*rows*3*D*D*B/us/1e6:.0f} TFLOP/s, {nbytes/us/1e3:.0f} GB/s of its {nbytes/1e6:.1f} MB", flush=True)
