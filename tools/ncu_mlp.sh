#!/bin/bash
# Source-level ncu captures (--set full) of the fused tower MLP kernels at one shape; raw + source pages exported to CSV on the box.
#     gpurun --timeout 600 -- 'bash tools/ncu_mlp.sh r2 8192,96,6'
set -u
tag=${1:-r2}; shape=${2:-8192,96,6}
out=gpurun_out; tmp=/tmp/ncu_mlp_$tag
mkdir -p $out $tmp
python tools/mlp_probe.py --shape $shape > $out/${tag}_mlp_plain.log 2>&1 || { tail -5 $out/${tag}_mlp_plain.log; exit 1; }
cat $out/${tag}_mlp_plain.log
for w in fwd bwd; do
  ncu --set full --import-source on --clock-control none -k regex:mlp_fused_kernel -s 3 -c 1 -o $tmp/mlp_$w -f \
      python tools/mlp_probe.py --shape $shape --which $w --reps 2 > $out/${tag}_ncu_mlp_$w.log 2>&1
  ncu -i $tmp/mlp_$w.ncu-rep --page source --csv > $out/${tag}_mlp_${w}_source.csv 2>/dev/null
  ncu -i $tmp/mlp_$w.ncu-rep --page raw --csv > $out/${tag}_mlp_${w}_raw.csv 2>/dev/null
  gzip -f $out/${tag}_mlp_${w}_source.csv
  ls -la $out/${tag}_mlp_${w}_*
done
