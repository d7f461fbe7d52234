#!/bin/bash
# The measurement sequence of one build, for ONE gpurun call (about 6 GPU-minutes):
#     gpurun --timeout 600 -- 'bash tools/profile_round.sh r2_v1 list'
#     gpurun --timeout 600 -- 'bash tools/profile_round.sh r2_v1 full regex:gemm_pair_kernel'
# 1. parity tests (a kernel that is fast and wrong is not done), 2. the headline bench line, 3. per-op steady-state times,
# 4. ONE ncu pass per call, only after the same command has run clean without ncu: `list` = the launch list of one eager
# evaluation, `full` = one --set full capture of the kernel named in $3.  Everything lands in gpurun_out/<tag>_*; copy what should
# be judged into profiles/.
set -u
tag=${1:-round}
mode=${2:-list}
kernel=${3:-regex:gemm_pair_kernel}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { tail -5 $out/${tag}_bench.err; exit 1; }
cut -c1-200 $out/${tag}_bench.json
python tools/profile_ops.py --reps 10 > $out/${tag}_ops.txt 2>&1; grep -E "==|TOTAL" $out/${tag}_ops.txt
python tools/profile_step.py --T 6 > $out/${tag}_step.log 2>&1 || { tail -5 $out/${tag}_step.log; exit 1; }
if [ "$mode" = full ]; then
  ncu --profile-from-start off --set full --clock-control none --import-source on -k "$kernel" -s 40 -c 1 -o $out/${tag}_kernel -f \
      python tools/profile_step.py --T 6 > $out/${tag}_ncu_kernel.log 2>&1
  ls -la $out/${tag}_kernel.ncu-rep 2>/dev/null
else
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
      python tools/profile_step.py --T 6 > $out/${tag}_ncu_step.log 2>&1
  wc -l $out/${tag}_launches.csv
fi
