"""What a GEMM -> GEMM launch boundary costs: two dependent launches enqueued behind a long memset, per-CTA globaltimer stamps
(entry, prologue done, pdl_wait returned, work done, about to exit) of both.
    python tools/gemm_boundary.py [--shape1 2048,4608,1152 --shape2 2048,1152,4608]"""
import argparse
import ctypes as C
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from vaevar_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--shape1", default="2048,4608,1152")
ap.add_argument("--shape2", default="2048,1152,4608")
a = ap.parse_args()
lib = _lib.load()
dev = "cuda:0"
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
M1, N1, K1 = map(int, a.shape1.split(","))
M2, N2, K2 = map(int, a.shape2.split(","))
dt = torch.float16
A1 = torch.randn(1, M1, K1, device=dev).to(dt); W1 = (torch.randn(1, N1, K1, device=dev) * 0.05).to(dt); b1 = torch.randn(1, N1, device=dev)
o1 = torch.empty(1, M1, N1, device=dev, dtype=dt); aux = torch.empty_like(o1)
W2 = (torch.randn(1, N2, K2, device=dev) * 0.05).to(dt); b2 = torch.randn(1, N2, device=dev); res = torch.randn(1, M2, N2, device=dev)
o2 = torch.empty(1, M2, N2, device=dev)
assert K2 == N1 and M1 == M2
g1 = (P(A1), P(W1), P(b1), None, None, P(o1), P(aux), M1, N1, K1, 1, 1 | 16, st)          # fc1: GELU + saved gelu'
g2 = (P(o1), P(W2), P(b2), P(res), P(o2), None, None, M2, N2, K2, 1, 0 | 16, st)           # fc2: + bias + residual, fp32 out
for _ in range(3):
    lib.vv_test_gemm(*g1); lib.vv_test_gemm(*g2)
t1 = torch.zeros(160 * 64, dtype=torch.int64, device=dev); t2 = torch.zeros_like(t1)
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
flush.zero_()                                    # ~150 us of GPU work: both launches are enqueued before it ends
lib.vv_debug_gemm_trace(P(t1)); lib.vv_test_gemm(*g1)
lib.vv_debug_gemm_trace(P(t2)); lib.vv_test_gemm(*g2)
lib.vv_debug_gemm_trace(None)
torch.cuda.synchronize()
T1 = t1.cpu().numpy().reshape(160, 64); T2 = t2.cpu().numpy().reshape(160, 64)
l1 = np.nonzero(T1[:, 11])[0]; l2 = np.nonzero(T2[:, 11])[0]
z = T1[l1, 11].min()
f = lambda T, l, k: (T[l, k].min() - z, int(np.median(T[l, k])) - z, T[l, k].max() - z)
names = {11: "entry", 12: "prologue done", 10: "pdl_wait returned", 13: "work done", 14: "about to exit"}
for tag, T, l in (("GEMM1", T1, l1), ("GEMM2", T2, l2)):
    print(f"== {tag}: {len(l)} CTAs; ns since GEMM1's first CTA entered (min / median / max over CTAs)")
    for k in (11, 12, 10, 13, 14):
        mn, md, mx = f(T, l, k)
        print(f"   {names[k]:18s} {mn:8d} {md:8d} {mx:8d}")
print("boundary: GEMM1 last 'work done' -> GEMM2 median 'pdl_wait returned' =", int(np.median(T2[l2, 10])) - T1[l1, 13].max(), "ns")
print("          GEMM1 last exit -> GEMM2 first entry =", T2[l2, 11].min() - T1[l1, 14].max(), "ns")
clk = lambda T, l, k: int(np.median(T[l, k] - T[l, 0]))
lead = [c for c in l2 if c % 2 == 0]
print("GEMM2 first k-block ready after pdl_wait (clk):", int(np.median(T2[lead, 40] - T2[lead, 0])), " work (clk):", clk(T2, l2, 1))
for tag, T, l in (("GEMM1", T1, l1), ("GEMM2", T2, l2)):
    ld = [c for c in l if c % 2 == 0]
    g = lambda k: int(np.median([T[c, k] - T[c, 0] for c in ld if T[c, k]])) if any(T[c, k] for c in ld) else -1
    print(f"{tag} (leader CTAs, clk after pdl_wait): producer done {g(1)}, MMA tile0 start {g(2)} commit {g(4)}, tile1 start {g(3)} commit {g(5)}; "
          f"epilogue warp0: tile0 acc seen {g(6)} done {g(8)}, tile1 acc seen {g(7)} done {g(9)}; exit-stamp (ns after wait) "
          f"{int(np.median(T[ld, 14] - T[ld, 10]))}")
