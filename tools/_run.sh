python -m pytest tests -m gpu -x -q -k "gemm or cost_small or net_small or full_T6" > gpurun_out/pytest_s.log 2>&1; tail -2 gpurun_out/pytest_s.log; grep -i "fail" gpurun_out/pytest_s.log | head
VV_GEMM_AUTOTUNE_LOG=1 python tools/profile_ops.py > gpurun_out/ops10.txt 2> gpurun_out/tune.log; grep -E "==|TOTAL" gpurun_out/ops10.txt; sort -u gpurun_out/tune.log | head -80
python bench.py --no-cpu-baseline > gpurun_out/bench10.log 2>gpurun_out/bench10.err; cut -c1-200 gpurun_out/bench10.log
