"""Summarise an `ncu --page source --csv` export (optionally .gz): stall reasons per code region (regions = runs of SASS with the
same execution count), the instructions that hold the most stall samples, and headline metrics from the matching raw page.
    python tools/ncu_stalls.py gpurun_out/r2_qkv_tower_source.csv.gz [gpurun_out/r2_qkv_tower_raw.csv]"""
import csv
import gzip
import sys

REASONS = ['stall_barrier', 'stall_branch_resolving', 'stall_dispatch', 'stall_drain', 'stall_lg', 'stall_long_sb', 'stall_math', 'stall_membar',
           'stall_mio', 'stall_misc', 'stall_no_inst', 'stall_not_selected', 'stall_selected', 'stall_short_sb', 'stall_sleep', 'stall_tex', 'stall_wait']


def main(src, raw=None, top=25):
    op = gzip.open if src.endswith(".gz") else open
    rows = list(csv.reader(op(src, "rt")))
    print(rows[0][1][:160])
    hdr, data = rows[1], rows[2:]
    i_s, i_src, i_ex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
    idx = {r: hdr.index(r) for r in REASONS}
    tot = sum(int(r[i_s]) for r in data)
    print(f"{len(data)} SASS instructions, {tot} stall samples")
    # regions by execution count
    regs, start = [], 0
    for k in range(1, len(data) + 1):
        if k == len(data) or data[k][i_ex] != data[start][i_ex]:
            regs.append((start, k)); start = k
    print("regions holding >= 3 % of the samples (start, end, executions, samples, top reasons):")
    for lo, hi in regs:
        s = sum(int(r[i_s]) for r in data[lo:hi])
        if s < 0.03 * tot:
            continue
        d = {q[6:]: sum(int(r[idx[q]] or 0) for r in data[lo:hi]) for q in REASONS}
        d = sorted(((v, k) for k, v in d.items() if v), reverse=True)[:5]
        print(f"  [{lo:5d},{hi:5d}) x{data[lo][i_ex]:>8s}  {s:6d} ({100 * s / tot:4.1f} %)  " + ", ".join(f"{k} {v}" for v, k in d))
    print(f"top {top} instructions:")
    for k in sorted(range(len(data)), key=lambda k: -int(data[k][i_s]))[:top]:
        r = data[k]
        d = {q[6:]: int(r[idx[q]]) for q in REASONS if r[idx[q]] not in ('', '0')}
        print(f"  {k:5d} {int(r[i_s]):5d} x{r[i_ex]:>8s}  {r[i_src].strip()[:70]:70s} {d}")
    if raw:
        rr = list(csv.reader(open(raw)))
        h, v = rr[0], rr[2]
        want = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_tensor", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
                "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread"]
        for w in want:
            for j, name in enumerate(h):
                if name.startswith(w):
                    print(f"  {name} [{rr[1][j]}] = {v[j]}")


if __name__ == "__main__":
    main(*sys.argv[1:3])
