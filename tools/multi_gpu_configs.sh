#!/bin/bash
# BASELINE.json configs[3] and configs[4] on N GPUs of one node (one process per GPU, NCCL only for the final reductions):
#     gpurun --gpus 8 --timeout 900 -- 'bash tools/multi_gpu_configs.sh 8'
set -u
n=${1:-8}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
$TR --master-port 29611 tools/run_cases.py --cases 64 --check profiles/r2_cases_64_1gpu.json --out $out/r2_cases_64_${n}gpu.json > $out/r2_cases_${n}gpu.log 2>&1
echo "cases rc=$?"; tail -c 700 $out/r2_cases_${n}gpu.log
if [ "${2:-cycles}" = cycles ]; then
  $TR --master-port 29612 tools/run_cycles.py --cycles 30 --out $out/r2_cycles_30_${n}gpu.json > $out/r2_cycles_${n}gpu.log 2>&1
  echo "cycles rc=$?"; tail -c 1200 $out/r2_cycles_${n}gpu.log
fi
