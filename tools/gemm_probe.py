"""Run one GEMM configuration of the engine a few times (for ncu captures and event timing of a single shape / epilogue).
    python tools/gemm_probe.py --shape 2048,4608,1152,1 --variant gelu_aux [--reps 20]
Variants: bf16 | f32 | f32_res | f32_bf16_res | gelu | gelu_aux | dgelu   (the epilogues the engine's plans use)."""
import argparse
import ctypes as C
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch

from vaevar_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="2048,4608,1152,1")
ap.add_argument("--variant", default="gelu_aux")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
M, N, K, B = map(int, a.shape.split(","))
lib = _lib.load()
dev = "cuda:0"
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
A = torch.randn(B, M, K, device=dev).bfloat16(); W = (torch.randn(B, N, K, device=dev) * 0.05).bfloat16()
bias = torch.randn(B, N, device=dev); res = torch.randn(B, M, N, device=dev)
of = torch.empty(B, M, N, device=dev); ob = torch.empty(B, M, N, device=dev, dtype=torch.bfloat16)
aux = torch.randn(B, M, N, device=dev).bfloat16()
variants = {"bf16": (None, None, None, ob, None, 0), "f32": (None, None, of, None, None, 0),
            "f32_res": (bias, res, of, None, None, 0), "f32_bf16_res": (bias, res, of, ob, None, 0),
            "gelu": (bias, None, None, ob, None, 1), "gelu_aux": (bias, None, None, ob, aux, 1),
            "dgelu": (None, None, None, ob, aux, 2)}
names = list(variants) if a.variant == "all" else a.variant.split(",")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in names:
    b_, r_, f_, o_, a_, epi = variants[name]
    args = (P(A), P(W), P(b_), P(r_), P(f_), P(o_), P(a_), M, N, K, B, epi, st)
    for _ in range(3):
        _lib.check(lib.vv_test_gemm(*args))
    torch.cuda.synchronize()
    ts = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(a.reps):
        flush.zero_()
        e0.record()
        lib.vv_test_gemm(*args)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{M}x{N}x{K}x{B} {name}: median {med*1e3:.1f} us (min {ts[0]*1e3:.1f}) = {2*M*N*K*B/med/1e9:.0f} TFLOP/s, L2 flushed", flush=True)
