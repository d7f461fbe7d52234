#!/bin/bash
# Round-2 profiling pass over ONE eager decoder evaluation (T=1, 271 launches): a sections-level capture of every launch
# (exported to CSV on the box -- the .ncu-rep files are too large to travel) and source-level captures of two kernels.
#     gpurun --timeout 900 -- 'bash tools/ncu_round2.sh r2_base'
set -u
tag=${1:-r2}
out=gpurun_out
tmp=/tmp/ncu_$tag
mkdir -p $out $tmp
python tools/profile_step.py --T 1 > $out/${tag}_plain.log 2>&1 || { tail -5 $out/${tag}_plain.log; exit 1; }
tail -1 $out/${tag}_plain.log
SEC="--section SpeedOfLight --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section InstructionStats"
ncu --profile-from-start off $SEC --clock-control none -o $tmp/all -f python tools/profile_step.py --T 1 > $out/${tag}_ncu_all.log 2>&1
tail -1 $out/${tag}_ncu_all.log
ncu -i $tmp/all.ncu-rep --page raw --csv > $out/${tag}_all_raw.csv 2>/dev/null; wc -c $out/${tag}_all_raw.csv
# source-level: the batched tower fc1 GEMM (LN-consume + GELU; 3rd GEMM launch) and the first trunk attention forward
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemm_pair_kernel -s 2 -c 1 -o $tmp/fc1 -f \
    python tools/profile_step.py --T 1 > $out/${tag}_ncu_fc1.log 2>&1
ncu -i $tmp/fc1.ncu-rep --page source --csv > $out/${tag}_fc1_source.csv 2>/dev/null
ncu -i $tmp/fc1.ncu-rep --page raw --csv > $out/${tag}_fc1_raw.csv 2>/dev/null
wc -c $out/${tag}_fc1_source.csv
gzip -f $out/${tag}_fc1_source.csv
ls -la $out/${tag}_*
