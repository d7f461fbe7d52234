"""Timeline of CTA 0 of one fused-MLP launch (clock64 stamps; vv_debug_mlp_trace): epilogue warp 0 and the MMA warp.
    python tools/mlp_trace.py --shape 8192,96,6 --which fwd"""
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
ap.add_argument("--which", default="fwd")
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
for name in a.which.split(","):
    fn = fns[name]
    fns["fwd"]()
    for _ in range(3):
        _lib.check(fn())
    trace = torch.zeros(128, dtype=torch.int64, device=dev)
    _lib.check(lib.vv_debug_mlp_trace(P(trace)))
    torch.cuda.synchronize()
    _lib.check(fn())
    torch.cuda.synchronize()
    _lib.check(lib.vv_debug_mlp_trace(None))
    t = trace.cpu().numpy()
    t0 = min(int(v) for v in t if v > 0)
    print(f"== mlp {name} {rows}x{D}x{B}: CTA 0, cycles since its first stamp")
    print("epilogue warp 0:", " ".join(str(int(v) - t0) for v in t[:64] if v > 0))
    print("MMA warp:       ", " ".join(str(int(v) - t0) for v in t[64:] if v > 0))
