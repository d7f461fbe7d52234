#!/bin/bash
# compute-sanitizer over the small configurations: memcheck over one T=3 cost+grad of the shrunken networks
# (eager launches, every kernel family of the engine) and one LGUnet_all_1 application at the mid-size configuration; racecheck and
# synccheck over the same cost+grad.  Results: gpurun_out/<tag>_sanitize_*.log, summary lines on stdout.
#     gpurun --timeout 1500 -- 'bash tools/sanitize.sh r2'
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
cat > /tmp/san_cost.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from vaevar_b200.config import DECODER_FULL, FLOW_FULL, small
from vaevar_b200.engine import Engine
from vaevar_b200.synth import make_case, make_state_dict
ds, fs = small(DECODER_FULL), small(FLOW_FULL)
e = Engine(ds, fs, T=3, use_graph=False)
e.load_state_dict(0, make_state_dict(ds, seed=0)); e.load_state_dict(1, make_state_dict(fs, seed=1)); e.finalize()
c = make_case(3, *ds.img_size, obs_frac=0.10, seed=0)
e.set_case(c["xb"], c["yo"], c["H"], c["R"], 1.0)
J, g = e.cost_grad(torch.from_numpy(c["z"]).cuda())
torch.cuda.synchronize()
print("cost+grad under the sanitizer: J=%.8g |g|=%.6g launches=%d" % (float(J[0]), float(g.norm()), e.last_launch_count))
PY
cat > /tmp/san_net1.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from vaevar_b200.config import FORECAST_MID
from vaevar_b200.forecast import ForecastNet
from vaevar_b200.synth import make_state_dict_net1
n = ForecastNet(FORECAST_MID, keep_out=69)
n.load_state_dict(make_state_dict_net1(FORECAST_MID, seed=11, rich=True)); n.finalize()
y = n.forward(torch.randn(69, *FORECAST_MID.img_size).cuda())
torch.cuda.synchronize()
print("LGUnet_all_1 under the sanitizer: |y|=%.6g launches=%d" % (float(y.norm()), n.last_launch_count))
PY
run() {  # name tool script limit
  timeout $4 compute-sanitizer --tool $2 --print-limit 20 --error-exitcode 9 python $3 > $out/${tag}_sanitize_$1.log 2>&1
  rc=$?
  echo "[$1] rc=$rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|under the sanitizer' $out/${tag}_sanitize_$1.log | tr '\n' ' ')"
}
run memcheck_cost memcheck /tmp/san_cost.py 600
run memcheck_net1 memcheck /tmp/san_net1.py 600
run racecheck_cost racecheck /tmp/san_cost.py 900
run synccheck_cost synccheck /tmp/san_cost.py 600
