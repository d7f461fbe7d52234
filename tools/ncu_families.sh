#!/bin/bash
# Source-level ncu captures (--set full, one launch each) of three GEMM families of ONE eager decoder evaluation (T=1):
#   qkv_tower  gemm #0   8192x288x96 x6   LN-consume, linear        (tower family: epilogue-bound)
#   proj_trunk gemm #19  2048x1152x1152   statistics producer + res (N=1152 family, K=1152)
#   fc2_trunk  gemm #21  2048x1152x4608   statistics producer + res (N=1152 family, K=4608)
# The .ncu-rep files are too large to travel: raw + source pages are exported to CSV on the box.
#     gpurun --timeout 900 -- 'bash tools/ncu_families.sh r2'
set -u
tag=${1:-r2}
out=gpurun_out
tmp=/tmp/ncu_$tag
mkdir -p $out $tmp
python tools/profile_step.py --T 1 > $out/${tag}_fam_plain.log 2>&1 || { tail -5 $out/${tag}_fam_plain.log; exit 1; }
tail -1 $out/${tag}_fam_plain.log
for spec in qkv_tower:0 proj_trunk:19 fc2_trunk:21; do
  name=${spec%%:*}; skip=${spec##*:}
  ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemm_pair_kernel -s $skip -c 1 -o $tmp/$name -f \
      python tools/profile_step.py --T 1 > $out/${tag}_ncu_$name.log 2>&1
  ncu -i $tmp/$name.ncu-rep --page source --csv > $out/${tag}_${name}_source.csv 2>/dev/null
  ncu -i $tmp/$name.ncu-rep --page raw --csv > $out/${tag}_${name}_raw.csv 2>/dev/null
  gzip -f $out/${tag}_${name}_source.csv
  ls -la $out/${tag}_${name}_*
done
