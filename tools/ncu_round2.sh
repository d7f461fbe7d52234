#!/bin/bash
# Round-2 profiling pass: ncu --set full captures of the kernels VERDICT r1 names (tower GEMMs, N=1152 trunk GEMMs, window
# attention, LayerNorm backward) out of ONE eager decoder evaluation (T=1).  Each capture only after the plain command ran clean.
#     gpurun --timeout 900 -- 'bash tools/ncu_round2.sh r2_base'
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
python tools/profile_step.py --T 1 > $out/${tag}_plain.log 2>&1 || { tail -5 $out/${tag}_plain.log; exit 1; }
tail -1 $out/${tag}_plain.log
N="ncu --profile-from-start off --set full --clock-control none"
$N -k regex:gemm_pair_kernel -c 26 -o $out/${tag}_gemm_fwd -f python tools/profile_step.py --T 1 > $out/${tag}_ncu1.log 2>&1; tail -1 $out/${tag}_ncu1.log
$N --import-source on -k regex:gemm_pair_kernel -c 4 -o $out/${tag}_gemm_fwd_src -f python tools/profile_step.py --T 1 > $out/${tag}_ncu1s.log 2>&1
$N -k regex:gemm_pair_kernel -s 87 -c 14 -o $out/${tag}_gemm_bwd -f python tools/profile_step.py --T 1 > $out/${tag}_ncu2.log 2>&1
$N --import-source on -k "regex:attn_kernel|ln_bwd_kernel|ln_fwd_kernel|p2t_kernel|t2p_kernel" -c 14 -o $out/${tag}_misc_fwd -f python tools/profile_step.py --T 1 > $out/${tag}_ncu3.log 2>&1
$N --import-source on -k "regex:attn_kernel|ln_bwd_kernel|ln_fwd_kernel|p2t_kernel|t2p_kernel" -s 30 -c 16 -o $out/${tag}_misc_bwd -f python tools/profile_step.py --T 1 > $out/${tag}_ncu4.log 2>&1
ls -la $out/${tag}_*.ncu-rep
