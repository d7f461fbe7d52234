"""Per-CTA timeline of one GEMM launch (clock64 stamps written by the kernel when vv_debug_gemm_trace is armed).
    python tools/gemm_trace.py --shape 2048,4608,1152,1 --variant gelu_aux [--bn 256] [--flush 1]
Slots per CTA (uint64): 0 start, 1 end, 2+ti MMA tile start, 4+ti MMA tile committed, 6+ti epilogue sees accumulator,
8+ti epilogue done (ti < 2), 10 globaltimer at start, 16+it producer issues k-block it, 40+it MMA sees k-block it (it < 24)."""
import argparse
import ctypes as C
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from vaevar_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="2048,4608,1152,1")
ap.add_argument("--variant", default="gelu_aux")
ap.add_argument("--bn", type=int, default=0)
ap.add_argument("--flush", type=int, default=1)
ap.add_argument("--f16", type=int, default=0)
ap.add_argument("--mode", type=int, default=0)
a = ap.parse_args()
if a.bn:
    os.environ["VV_GEMM_BN"] = str(a.bn)
M, N, K, B = map(int, a.shape.split(","))
lib = _lib.load()
dev = "cuda:0"
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
dt = torch.float16 if a.f16 else torch.bfloat16
A = torch.randn(B, M, K, device=dev).to(dt); W = (torch.randn(B, N, K, device=dev) * 0.05).to(dt)
bias = torch.randn(B, N, device=dev); res = torch.randn(B, M, N, device=dev)
of = torch.empty(B, M, N, device=dev); ob = torch.empty(B, M, N, device=dev, dtype=dt)
aux = torch.randn(B, M, N, device=dev).to(dt)
variants = {"bf16": (None, None, None, ob, None, 0), "f32": (None, None, of, None, None, 0),
            "f32_res": (bias, res, of, None, None, 0), "f32_bf16_res": (bias, res, of, ob, None, 0),
            "gelu": (bias, None, None, ob, None, 1), "gelu_aux": (bias, None, None, ob, aux, 1),
            "dgelu": (None, None, None, ob, aux, 2)}
b_, r_, f_, o_, a_, epi = variants[a.variant]
args = (P(A), P(W), P(b_), P(r_), P(f_), P(o_), P(a_), M, N, K, B, epi | (16 if a.f16 else 0), st)
_lib.check(lib.vv_debug_gemm_mode(a.mode))
for _ in range(3):
    _lib.check(lib.vv_test_gemm(*args))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    lib.vv_test_gemm(*args)
e1.record(); torch.cuda.synchronize()
print(f"back-to-back x20 (no flush): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
trace = torch.zeros(160 * 64, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_lib.check(lib.vv_debug_gemm_trace(P(trace)))
if a.flush:
    flush.zero_()
torch.cuda.synchronize()
_lib.check(lib.vv_test_gemm(*args))
torch.cuda.synchronize()
_lib.check(lib.vv_debug_gemm_trace(None))
t = trace.cpu().numpy().reshape(160, 64)
live = np.nonzero(t[:, 0])[0]
print(f"== {M}x{N}x{K}x{B} {a.variant} bn={a.bn or 'auto'} flush={a.flush} mode={a.mode}: {len(live)} CTAs traced")
g0 = t[live, 10].min()
print("globaltimer start spread (ns): max-min =", int(t[live, 10].max() - g0))
rel = lambda c, k: (t[c, k] - t[c, 0]) if t[c, k] else -1
dur = np.array([rel(c, 1) for c in live])
print(f"CTA duration (clk): min {dur.min()} median {int(np.median(dur))} max {dur.max()}")
for c in [live[0], live[len(live) // 2 & ~1], live[-2]]:
    print(f"-- CTA {c} (leader): end {rel(c,1)}")
    print("   MMA tile start", [int(rel(c, 2 + i)) for i in range(2)], "commit", [int(rel(c, 4 + i)) for i in range(2)])
    print("   EPI acc seen  ", [int(rel(c, 6 + i)) for i in range(2)], "done  ", [int(rel(c, 8 + i)) for i in range(2)])
    pi = [int(rel(c, 16 + i)) for i in range(24)]
    mr = [int(rel(c, 40 + i)) for i in range(24)]
    print("   producer issue:", pi)
    print("   mma ready     :", mr)
    print("   load latency  :", [m - p_ if m >= 0 and p_ >= 0 else -1 for p_, m in zip(pi, mr)])
# aggregate over leader CTAs: steady-state interval between consecutive k-blocks at the MMA
lead = [c for c in live if c % 2 == 0]
iv = np.array([[t[c, 40 + i + 1] - t[c, 40 + i] for i in range(8, 22)] for c in lead if t[c, 40 + 22]])
if len(iv):
    print(f"MMA k-block interval (clk), k-blocks 8..22: mean {iv.mean():.0f} median {np.median(iv):.0f} p90 {np.percentile(iv, 90):.0f}")
lat = np.array([[t[c, 40 + i] - t[c, 16 + i] for i in range(8, 22)] for c in lead if t[c, 40 + 22]])
if len(lat):
    print(f"issue->ready latency (clk), k-blocks 8..22: mean {lat.mean():.0f} median {np.median(lat):.0f} p90 {np.percentile(lat, 90):.0f}")
first = np.array([t[c, 40] - t[c, 0] for c in lead])
print(f"first k-block ready after start (clk): median {int(np.median(first))} max {first.max()}")
