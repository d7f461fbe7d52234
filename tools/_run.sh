python -m pytest tests -m gpu -x -q -k "attn or net_small or cost_small or full_T6" > gpurun_out/pytest_s.log 2>&1; tail -2 gpurun_out/pytest_s.log; grep -i "fail" gpurun_out/pytest_s.log | head
python tools/profile_ops.py > gpurun_out/ops8.txt 2>&1; grep -E "==|TOTAL|attn" gpurun_out/ops8.txt
python bench.py --no-cpu-baseline > gpurun_out/bench8.log 2>gpurun_out/bench8.err; cut -c1-200 gpurun_out/bench8.log
