"""Steady-state per-launch times of one decoder and one flow application (forward and backward plans), measured with
CUDA events by the engine itself (vv_profile_ops): no profiler, warm caches, back-to-back launches.
    python tools/profile_ops.py [--reps 20] > profiles/ops.txt"""
import argparse
import collections
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import build_inputs
from vaevar_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dcfg, fcfg, sd_d, sd_f, case = build_inputs(2, 0.10, 0)
eng = Engine(dcfg, fcfg, T=2, use_graph=False)
eng.load_state_dict(0, sd_d); eng.load_state_dict(1, sd_f); eng.finalize()
eng.set_case(case["xb"], case["yo"], case["H"], case["R"], 1.0)
grand = 0.0
for app, name in ((0, "decoder"), (1, "flow")):
    for bwd in (False, True):
        ops = eng.profile_ops(app, bwd, a.reps)
        tot = sum(o["ms"] for o in ops)
        grand += tot
        agg = collections.OrderedDict()
        for o in ops:
            key = (o["kind"],) + tuple(o["shape"])
            t = agg.setdefault(key, [0, 0.0, 0.0])
            t[0] += 1; t[1] += o["ms"]; t[2] += o["flop"]
        print(f"== {name} {'backward' if bwd else 'forward'}: {len(ops)} launches, {tot:.3f} ms (sum of per-op steady-state times)")
        for key, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            tf = f"{fl/ms/1e9:8.0f} TFLOP/s" if fl else " " * 15
            print(f"   {ms*1e3:9.1f} us {100*ms/tot:5.1f}%  n={n:3d} avg={ms/n*1e3:7.1f} us {tf}  {key[0]:9s} shape={key[1:]}")
print(f"TOTAL decoder+flow fwd+bwd: {grand:.3f} ms -> T=6 estimate {grand/2*6:.1f} ms")
