"""ms per T=6 cost+grad evaluation (CUDA graph, 20 timed evaluations after 4 warm-ups) and J, for the library VV_LIB points at --
the A/B harness of kernel experiments (tools/build_variant.sh)."""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch

from bench import build_inputs
from vaevar_b200 import _lib
from vaevar_b200.engine import Engine

T = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dcfg, fcfg, sd_d, sd_f, case = build_inputs(T, 0.10, 0)
eng = Engine(dcfg, fcfg, T=T)
eng.load_state_dict(0, sd_d); eng.load_state_dict(1, sd_f); eng.finalize()
eng.set_case(case["xb"], case["yo"], case["H"], case["R"], 1.0)
z = torch.from_numpy(case["z"]).cuda()
for _ in range(4):
    J, g = eng.cost_grad(z)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        J, g = eng.cost_grad(z)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
fam = {}
for app, mult in ((0, 1), (1, T - 1)):
    for bwd in (False, True):
        for o in eng.profile_ops(app, bwd, 5):
            k = o["kind"] if o["kind"] != "gemm" else ("gemm trunk" if o["shape"][3] == 1 else "gemm towers")
            fam[k] = fam.get(k, 0.0) + o["ms"] * mult
print(f"{_lib.LIB_PATH.name}: {best:.3f} ms per cost+grad, J={float(J[0]):.9g} |g|={float(g.double().norm()):.7g}; per-step family ms: " +
      ", ".join(f"{k} {v:.2f}" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])))
