python -m pytest tests -m gpu -x -q -k "gemm or net_small or cost_small" > gpurun_out/pytest_small.log 2>&1; tail -3 gpurun_out/pytest_small.log
python tools/gemm_sweep.py --bns 0,128,192,256 --variants bf16,f32_res,gelu_aux --shapes "2048,4608,1152,1;2048,1152,4608,1;2048,3456,1152,1;2048,1152,3456,1;2048,1152,1152,1;2048,4608,9216,1" > gpurun_out/sweep_flush2.txt 2>&1
cat gpurun_out/sweep_flush2.txt
for cfg in "2048,4608,1152,1 gelu_aux 256 1 0" "2048,1152,4608,1 f32_res 128 1 0" "2048,4608,9216,1 bf16 256 0 0" "2048,4608,9216,1 bf16 128 0 0" "2048,4608,9216,1 bf16 256 0 1" "2048,4608,9216,1 bf16 128 0 1" ; do
  set -- $cfg
  python tools/gemm_trace.py --shape $1 --variant $2 --bn $3 --flush $4 --mode $5
done > gpurun_out/trace2.txt 2>&1
grep -E "==|back-to|interval|latency \(clk|CTA duration" gpurun_out/trace2.txt
