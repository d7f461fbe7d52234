"""One eager cost+grad evaluation bracketed by cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/profile_step.py --T 6
(and `--set full -k regex:gemm_tn` captures).  Without ncu it just runs and prints the event-timed step."""
import argparse
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch

from bench import build_inputs
from vaevar_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=6)
ap.add_argument("--recompute", type=int, default=0)
a = ap.parse_args()
dcfg, fcfg, sd_d, sd_f, case = build_inputs(a.T, 0.10, 0)
eng = Engine(dcfg, fcfg if a.T > 1 else None, T=a.T, recompute=bool(a.recompute), use_graph=False)
eng.load_state_dict(0, sd_d)
if a.T > 1:
    eng.load_state_dict(1, sd_f)
eng.finalize()
eng.set_case(case["xb"], case["yo"], case["H"], case["R"], 1.0)
z = torch.from_numpy(case["z"]).cuda()
for _ in range(2):
    J, g = eng.cost_grad(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
e0.record()
J, g = eng.cost_grad(z)
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"T={a.T} eager cost+grad: {e0.elapsed_time(e1):.3f} ms, launches={eng.last_launch_count}, J={float(J[0]):.8g}")
