"""Sweep the engine's GEMM over shapes / epilogues / tile widths, with torch.matmul (cuBLAS) on the same shape beside it.
    python tools/gemm_sweep.py [--reps 15] [--flush 0|1] [--shapes M,N,K,B;...] [--bns 0,64,128,192,256] [--variants bf16,...]
BN = 0 lets the dispatcher pick.  Times are CUDA-event medians of single launches (L2 flushed between launches when --flush 1,
otherwise the operands stay L2-resident the way activations are inside the engine)."""
import argparse
import ctypes as C
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch

from vaevar_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="2048,4608,1152,1;2048,1152,4608,1;2048,3456,1152,1;2048,1152,3456,1;2048,1152,1152,1")
ap.add_argument("--variants", default="bf16,f32_res,gelu_aux,dgelu")
ap.add_argument("--bns", default="0")
ap.add_argument("--reps", type=int, default=15)
ap.add_argument("--flush", type=int, default=1)
ap.add_argument("--f16", type=int, default=0)
ap.add_argument("--cublas", type=int, default=1)
a = ap.parse_args()
lib = _lib.load()
dev = "cuda:0"
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        if a.flush:
            flush.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for shp in a.shapes.split(";"):
    M, N, K, B = map(int, shp.split(","))
    dt = torch.float16 if a.f16 else torch.bfloat16
    A = torch.randn(B, M, K, device=dev).to(dt); W = (torch.randn(B, N, K, device=dev) * 0.05).to(dt)
    bias = torch.randn(B, N, device=dev); res = torch.randn(B, M, N, device=dev)
    of = torch.empty(B, M, N, device=dev); ob = torch.empty(B, M, N, device=dev, dtype=dt)
    aux = torch.randn(B, M, N, device=dev).to(dt)
    variants = {"bf16": (None, None, None, ob, None, 0), "f32": (None, None, of, None, None, 0),
                "f32_res": (bias, res, of, None, None, 0), "f32_bf16_res": (bias, res, of, ob, None, 0),
                "gelu": (bias, None, None, ob, None, 1), "gelu_aux": (bias, None, None, ob, aux, 1),
                "dgelu": (None, None, None, ob, aux, 2)}
    fl = 2.0 * M * N * K * B
    if a.cublas:
        Wt = W.transpose(1, 2)
        med, mn = timeit(lambda: torch.matmul(A, Wt, out=ob))
        print(f"{M}x{N}x{K}x{B} cuBLAS {str(dt)[6:]:8s}          : median {med*1e3:7.1f} us (min {mn*1e3:7.1f}) = {fl/med/1e9:6.0f} TFLOP/s", flush=True)
    for bn in a.bns.split(","):
        if int(bn) > 0:
            os.environ["VV_GEMM_BN"] = bn
        else:
            os.environ.pop("VV_GEMM_BN", None)
        for name in a.variants.split(","):
            if name.startswith("ln_"):                      # folded-LayerNorm consumer (statistics with 2 partials per row)
                epi = 1 if "gelu" in name else 0
                stats = torch.rand(B * 2 * M * 2, device=dev) + 1.0
                colsum = torch.randn(B, N, device=dev)
                largs = (P(A), P(W), P(bias), P(colsum), P(stats), 2, K, 1e-5, P(ob), P(aux) if "aux" in name else None, None, None, None,
                         M, N, K, B, epi | (16 if a.f16 else 0), st)
                _lib.check(lib.vv_test_gemm_ln(*largs))
                med, mn = timeit(lambda: lib.vv_test_gemm_ln(*largs))
                print(f"{M}x{N}x{K}x{B} BN={bn:>3s} {name:14s}: median {med*1e3:7.1f} us (min {mn*1e3:7.1f}) = {fl/med/1e9:6.0f} TFLOP/s", flush=True)
                continue
            b_, r_, f_, o_, a_, epi = variants[name]
            args = (P(A), P(W), P(b_), P(r_), P(f_), P(o_), P(a_), M, N, K, B, epi | (16 if a.f16 else 0), st)
            _lib.check(lib.vv_test_gemm(*args))
            med, mn = timeit(lambda: lib.vv_test_gemm(*args))
            print(f"{M}x{N}x{K}x{B} BN={bn:>3s} {name:14s}: median {med*1e3:7.1f} us (min {mn*1e3:7.1f}) = {fl/med/1e9:6.0f} TFLOP/s", flush=True)
    del A, W, bias, res, of, ob, aux
os.environ.pop("VV_GEMM_BN", None)
