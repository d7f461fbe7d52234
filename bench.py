"""Headline benchmark: one 4D-Var cost+gradient evaluation J(z), grad_z J (da_4dvar.py:1183-1208, 1242-1246).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--T 6] [--obs-frac 0.1]

Workload (BASELINE.json configs[1]): 6-step window, 69x128x256 state, VAE decoder + 5 applications of the flow
model (both 216 M-parameter U-shaped Swin networks, random init), 10 % column observations, one case per GPU.
A "step" is one closure() = cost + gradient.  N > 1: one process per GPU (torchrun), independent cases, no
data-path collective (replicas), only the timing reduction goes through NCCL.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# Algorithmic work per cost+grad: forward + input-gradient, no weight gradients, no recompute.
GMAC_DEC, GMAC_FLOW = 446.05, 446.17   # flow: discarded log-var half of the final projection not counted


def algorithmic_tflop(T: int) -> float:
    return 2 * 2 * (GMAC_DEC + (T - 1) * GMAC_FLOW) * 1e9 / 1e12


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), d["hbm_gbs"], "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(T, obs_frac, seed):
    from vaevar_b200.config import DECODER_FULL, FLOW_FULL
    from vaevar_b200.synth import make_case, make_state_dict
    sd_d = make_state_dict(DECODER_FULL, seed=0)
    sd_f = make_state_dict(FLOW_FULL, seed=1) if T > 1 else None
    case = make_case(T, 128, 256, obs_frac=obs_frac, seed=seed)
    return DECODER_FULL, FLOW_FULL, sd_d, sd_f, case


def workload_config(T, obs_frac):
    """The `config` object BOTH arms print, key for key (the reference arm runs the same workload on the host cores)."""
    return {"workload": f"4D-Var {T}-step window cost+grad (VAE decoder + {T-1} flow-model applications, forward + adjoint), "
                        f"1 case per GPU, 69x128x256 state, {int(obs_frac*100)}% column obs (BASELINE.json configs[1])",
            "T": T, "obs_frac": obs_frac, "n_obs": int(69 * T * int(obs_frac * 128 * 256)), "state": [69, 128, 256], "latent": [32, 128, 256],
            "l2": "inputs >> L2: per-eval working set (2 x 0.86 GB 16-bit weights + ~1.3 GB stash per application) vs 126 MB L2; no flush needed"}


def cpu_oracle_eval(T, obs_frac, seed, repeats=1, budget_s=180.0, as_is=False, device="cpu", tf32=False):
    """Times the oracle (oracle/: reference algorithm, fp32 torch, all host threads) on closure() calls.
    as_is: weights carry requires_grad=True as in the reference (da_4dvar.py:590-603 never freezes them), so backward() also fills
    216 M weight gradients per network that nobody reads.  device="cuda": the same restatement run eagerly on the GPU (informational)."""
    import torch
    from oracle import cost as oc
    from oracle.lgunet import to_torch
    torch.set_num_threads(os.cpu_count() or 1)
    dcfg, fcfg, sd_d, sd_f, case = build_inputs(T, obs_frac, seed)
    prep = lambda sd: {k: v.to(device).requires_grad_(as_is and v.is_floating_point()) for k, v in to_torch(sd).items()}
    nets = oc.OracleNets(prep(sd_d), dcfg, prep(sd_f) if sd_f else None, fcfg)
    c = oc.Case(case).to(device)
    if device != "cpu":
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
    times, t_start = [], time.time()
    J = None
    for _ in range(repeats):
        if as_is:
            for sd in (nets.sd_dec, nets.sd_flow or {}):
                for v in sd.values():
                    v.grad = None                                   # optimizer.zero_grad(set_to_none) of the reference closure
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.time()
        J, _, _, _ = oc.cost_and_grad(case["z"], c, nets)
        if device != "cpu":
            torch.cuda.synchronize()
        times.append(time.time() - t0)
        if time.time() - t_start > budget_s:
            break
    return times, J, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference algorithm's CPU path (oracle port; the reference is pure PyTorch and its
    driver cannot be imported: da_4dvar.py:18,23-25) on this box's host cores, same workload and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Ts = args.T
    times, J, threads = cpu_oracle_eval(Ts, args.obs_frac, 0, repeats=args.warmup + args.steps, budget_s=240.0)
    timed = times[min(args.warmup, max(len(times) - 1, 0)):] or times
    ms = 1e3 * sum(timed) / len(timed)
    sample = (f"{len(timed)} full closure() calls (T={Ts}, 69x128x256, {int(args.obs_frac*100)}% obs) after {len(times)-len(timed)} warm-up; fp32 torch CPU, "
              f"weight gradients off (the as-is variant is in the b200 arm's cpu_baseline_as_is); rank 0 only at any --gpus N (per-replica metric)")
    line = {"impl": "reference", "metric": "ms per 4D-Var cost+grad eval (69x128x256)", "value": ms, "unit": "ms",
            "n_gpus": args.gpus, "steps": len(timed), "warmup": len(times) - len(timed), "ms_per_step": ms,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(Ts, args.obs_frac),
            "cpu_baseline": {"value": ms, "unit": "ms", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "J": J}
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """ONE JSON line on stdout: everything any library prints to file descriptor 1 during the run (NCCL's version banner, ...)
    is sent to stderr; emit() writes the result line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--T", type=int, default=6)
    ap.add_argument("--obs-frac", type=float, default=0.10)
    ap.add_argument("--recompute", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inflight", action="store_true", help="skip the several-cases-in-flight-per-GPU measurement")
    ap.add_argument("--no-forecast", action="store_true", help="skip the LGUnet_all_1 (721x1440 forecast network) measurement")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from vaevar_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    T = args.T
    dcfg, fcfg, sd_d, sd_f, case = build_inputs(T, args.obs_frac, seed=rank)     # one independent case per GPU
    eng = Engine(dcfg, fcfg if T > 1 else None, T=T, recompute=bool(args.recompute), use_graph=not args.no_graph, device=f"cuda:{local}")
    eng.load_state_dict(0, sd_d)
    if T > 1:
        eng.load_state_dict(1, sd_f)
    eng.finalize()
    eng.set_case(case["xb"], case["yo"], case["H"], case["R"], 1.0)
    z = torch.from_numpy(case["z"]).to(dev)
    Jb = torch.empty(3, dtype=torch.float64, device=dev)
    gb = torch.empty_like(z)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") --------------------------------------------------------------
    for _ in range(args.warmup):
        eng.cost_grad(z, Jb, gb)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        eng.cost_grad(z, Jb, gb)
    e1.record()
    barrier()
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.last_launch_count

    # ---- end-to-end through the public API with HOST buffers ("e2e") -----------------------------------
    z_host = torch.from_numpy(case["z"]).pin_memory()
    g_host = torch.empty_like(z_host).pin_memory()
    J_host = torch.empty(3, dtype=torch.float64).pin_memory()
    for _ in range(2):
        z.copy_(z_host, non_blocking=True); eng.cost_grad(z, Jb, gb); g_host.copy_(gb, non_blocking=True); J_host.copy_(Jb, non_blocking=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        z.copy_(z_host, non_blocking=True)
        eng.cost_grad(z, Jb, gb)
        g_host.copy_(gb, non_blocking=True)
        J_host.copy_(Jb, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller reads J and grad on the host every step
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2)

    # ---- "one or more cases per GPU" (SURVEY 8e): S independent cases in flight on this GPU, one engine + stream + host thread each;
    # aggregate cost+grad evaluations per second of the GPU.  The headline `value` stays the single-case latency. ----
    inflight = {"1": {"evals_per_s_per_gpu": 1e3 * args.steps / total_ms, "ms_per_eval_per_case": total_ms / args.steps}}
    extra_engs = []
    if T > 1 and not args.no_inflight:
        try:
            import threading
            for k in (1, 2):
                c2 = build_inputs(T, args.obs_frac, seed=1000 + 17 * rank + k)[4]
                e2 = Engine(dcfg, fcfg, T=T, recompute=bool(args.recompute), use_graph=not args.no_graph, device=f"cuda:{local}")
                e2.load_state_dict(0, sd_d); e2.load_state_dict(1, sd_f); e2.finalize()
                e2.set_case(c2["xb"], c2["yo"], c2["H"], c2["R"], 1.0)
                extra_engs.append((e2, torch.from_numpy(c2["z"]).to(dev)))
            pool = [(eng, z)] + extra_engs
            for S in (2, 3):
                gate = threading.Barrier(S + 1)
                errs = []

                def worker(e_, z_):
                    try:
                        torch.cuda.set_device(local)
                        with torch.cuda.stream(torch.cuda.Stream(device=dev)):
                            J_ = torch.empty(3, dtype=torch.float64, device=dev); g_ = torch.empty_like(z_)
                            for _ in range(args.warmup):
                                e_.cost_grad(z_, J_, g_)
                            torch.cuda.current_stream().synchronize()
                            gate.wait()
                            for _ in range(args.steps):
                                e_.cost_grad(z_, J_, g_)
                            torch.cuda.current_stream().synchronize()
                            gate.wait()
                    except Exception as ex:          # never leave the main thread waiting at the gate
                        errs.append(repr(ex)); gate.abort()

                th = [threading.Thread(target=worker, args=pool[i]) for i in range(S)]
                for t_ in th:
                    t_.start()
                gate.wait(); t0 = time.perf_counter()
                gate.wait(); dt = time.perf_counter() - t0
                for t_ in th:
                    t_.join()
                if errs:
                    raise RuntimeError(errs[0])
                inflight[str(S)] = {"evals_per_s_per_gpu": S * args.steps / dt, "ms_per_eval_per_case": 1e3 * dt / args.steps}
            torch.cuda.synchronize()
        except Exception as ex:
            inflight["error"] = repr(ex)
        for e2, _ in extra_engs:
            e2.close()
        extra_engs = []
        torch.cuda.empty_cache()

    # ---- one REAL analysis-forecast cycle through the cycle driver (da_4dvar.py:1314-1342 with the shipped script's Nit=4,
    # da_4dvar_script.sh:14): 4 x LBFGS.step(max_iter=10) + 5 diagnostic sweeps + the forecast step, every rank its own case ----
    import tempfile
    from vaevar_b200.cycle import CycledDA, TwinObs
    from vaevar_b200.da import VaeVar4D
    agent = VaeVar4D(dcfg, fcfg if T > 1 else None, None, None, da_win=T, Nit=4, device=f"cuda:{local}", verbose=False, engine=eng)
    cyc_s, cyc_evals = None, 0
    if T > 1:
        with tempfile.TemporaryDirectory() as tmp:
            run = CycledDA(agent, TwinObs(agent, torch.from_numpy(case["gt"][0]), obs_frac=args.obs_frac, seed=rank),
                           torch.from_numpy(case["xb"]), name=f"bench{rank}", root=tmp, n_cycles=2, resume=False)
            barrier()
            run.run_assimilation()              # cycle 0 warms up (lazy module loads, first L-BFGS / metric launches); cycle 1 is reported
            barrier()
        cyc = torch.tensor([run.cycle_seconds[-1]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(cyc, op=dist.ReduceOp.MAX)
        cyc_s = float(cyc)
        cyc_evals = int(sum(h["n_evals"] for h in agent.history[-4:]))       # as torch.optim.LBFGS counts them; 3 of them are served
                                                                             # from the optimiser's own state (vv_lbfgs_set_reuse)
        eng.set_case(case["xb"], case["yo"], case["H"], case["R"], 1.0)          # restore the benchmark case

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    import ctypes as C
    burst, sustained, hbm, src = peaks()
    tflop = algorithmic_tflop(T)
    ms_eval = total_ms / args.steps
    step_tfs = tflop / (ms_eval * 1e-3)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ---- roofline of the dominant kernel = the tcgen05 GEMM (gemm_pair_kernel, every Linear forward and input-gradient: ~75 % of the
    # step).  All its launches of one decoder and one flow application (forward + backward plans), each timed with CUDA events by the
    # engine's own op profiler (steady state, launches back to back as inside the step): achieved = sum of their algorithmic flops /
    # sum of their times, against the SUSTAINED measured bf16 peak (kernels timed inside a long step).  Per-family rows and the
    # whole-step figure beside it; ncu launch lists / captures of the same families: profiles/r2_*. ----
    fam, gemm_ms, gemm_fl, all_ms = {}, 0.0, 0.0, 0.0
    apps = ((0, 1), (1, T - 1)) if T > 1 else ((0, 1),)
    try:
        for app, mult in apps:
            for bwd in (False, True):
                for o in eng.profile_ops(app, bwd, 5):
                    all_ms += o["ms"] * mult
                    if o["kind"] not in ("gemm", "mlp_fwd", "mlp_bwd", "lin_fwd"):       # the tensor-core kernels: the tcgen05 GEMM and the fused tower MLP
                        continue
                    gemm_ms += o["ms"] * mult; gemm_fl += o["flop"] * mult
                    M_, N_, K_, B_ = o["shape"]
                    key = "towers: fused norm1 + qkv, and the MLP half of a block (norm2 + fc1 + GELU + fc2 + residual) / its input-VJP (mlp_fused_kernel, tcgen05 cta_group::1)" \
                          if o["kind"] != "gemm" else \
                          "trunk d=1152, N=1152 (proj, fc2, dgrads: 72 tiles on 74 SM pairs)" if (B_ == 1 and N_ == 1152) else \
                          "trunk d=1152, N>=3456 (qkv, fc1, dgrad of fc2)" if B_ == 1 else "towers d=96/192: qkv, proj, seams and their dgrads (batched over 6 variable groups)"
                    f_ = fam.setdefault(key, [0.0, 0.0, 0]); f_[0] += o["ms"] * mult; f_[1] += o["flop"] * mult; f_[2] += mult
    except Exception as ex:
        fam = {"error": [0.0, 0.0, repr(ex)]}
    gemm_tfs = gemm_fl / max(gemm_ms, 1e-9) / 1e9
    roof_rows = [{"family": k, "launches_per_step": v[2], "ms_per_step": round(v[0], 3), "share_of_gemm_time": round(v[0] / max(gemm_ms, 1e-9), 3),
                  "TFLOP/s": round(v[1] / max(v[0], 1e-9) / 1e9, 1), "frac_of_sustained": round(v[1] / max(v[0], 1e-9) / 1e9 / sustained, 3)}
                 for k, v in fam.items()]
    for r_ in roof_rows:                                      # dram__bytes_read + write of one d = 96 launch, `ncu --set full` (profiles/r2_ncu_mlp_*_summary.txt)
        if "mlp_fused_kernel" in r_["family"]:
            r_["dram_bytes_ncu"] = {"mlp forward 6x8192x96": 34.4e6, "mlp backward 6x8192x96": 99.6e6}
    # the trunk's fc1 shape alone, L2 flushed between launches (round 1's `roofline` row, kept as a second row)
    M, N, K = 2048, 4608, 1152
    A = torch.randn(1, M, K, device=dev).half(); W = (torch.randn(1, N, K, device=dev) * 0.05).half()
    bias = torch.randn(1, N, device=dev)
    ob = torch.empty(1, M, N, device=dev, dtype=torch.float16); aux = torch.empty_like(ob)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tk = []
    for i in range(25):
        flush.zero_()                                       # evict L2 (256 MiB > 126 MB) between timed launches
        e0.record()
        eng.lib.vv_test_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(bias.data_ptr()), None, None,
                             C.c_void_p(ob.data_ptr()), C.c_void_p(aux.data_ptr()), M, N, K, 1, 1 | 16, st)
        e1.record(); torch.cuda.synchronize()
        if i >= 5:
            tk.append(e0.elapsed_time(e1))
    k_tfs = 2.0 * M * N * K / (statistics.median(tk) * 1e-3) / 1e12
    roof_rows.append({"family": "fc1 2048x4608x1152 + bias + GELU + saved gelu' alone, L2 flushed (gemm_pair_kernel<256,4,fp16,LN_NONE,GELU>)",
                      "TFLOP/s": round(k_tfs, 1), "frac_of_burst": round(k_tfs / burst, 3), "dram_bytes_ncu": 16.24e6})

    # ---- HBM-bound kernels: achieved GB/s of their algorithmic bytes against the measured copy bandwidth.
    # Every timed launch starts on a COLD L2 (256 MiB overwritten before it, outside the events): an 85 MB working set re-run back to
    # back stays in the 126 MB L2 and rates above the HBM peak, which is not an HBM measurement. ----
    hbm_rows = []
    try:
        agg = {}
        for app, bwd in ((1 if T > 1 else 0, False), (1 if T > 1 else 0, True)):
            for o in eng.profile_ops(app, bwd, 4, flush_l2=True):
                k, sh = o["kind"], o["shape"]
                if k == "ln_bwd":
                    nbytes = sh[0] * sh[1] * max(sh[3], 1) * 16          # x, dres fp32 + dy bf16 read (10 B), dx fp32 + bf16 written (6 B)
                elif k in ("attn_fwd", "attn_bwd"):
                    # trunk: 2048 tokens x d=1152 (batch 1); towers: hd=32 appears for d=96 (8192 tokens) and d=192 (2048 tokens),
                    # 6 groups -- the profiler reports head_dim and batch only, so only the unambiguous trunk kernels are rated
                    if sh[1] != 192:
                        continue
                    nbytes = 2048 * 1152 * (8 if k == "attn_fwd" else 14)     # qkv (+dO) read, out / dqkv written, 2 B each
                else:
                    continue
                a_ = agg.setdefault((k,) + tuple(sh), [0, 0.0, nbytes])
                a_[0] += 1; a_[1] += o["ms"]
        for (k, *sh), (n, ms, nbytes) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            gbs = nbytes / (ms / n * 1e-3) / 1e9
            hbm_rows.append({"kernel": k, "shape": sh, "launches_per_application": n, "us_cold_l2": round(1e3 * ms / n, 2),
                             "algorithmic_MB": round(nbytes / 1e6, 2), "GB/s": round(gbs, 1), "frac": round(gbs / hbm, 3)})
        # the observation operator: fused gather + misfit over all T (16 B / observation) and its adjoint (12 B / observation)
        xn = torch.randn(T, 69, 128, 256, device=dev); Jo = torch.empty(1, dtype=torch.float64, device=dev); gx = torch.empty_like(xn)
        for with_adj in (False, True):
            ts = []
            for i in range(8):
                flush.zero_()
                e0.record()
                eng.lib.vv_test_obs(eng._h, C.c_void_p(xn.data_ptr()), C.c_void_p(Jo.data_ptr()), C.c_void_p(gx.data_ptr()) if with_adj else None, st)
                e1.record(); torch.cuda.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            us = 1e3 * statistics.median(ts)
            nbytes = eng.n_obs * (16 + 4) if not with_adj else eng.n_obs * (16 + 4 + 12) + xn.numel() * 4
            hbm_rows.append({"kernel": "obs_misfit + reduce" + (" + zero-fill + obs_adjoint x T" if with_adj else ""), "shape": [int(eng.n_obs)],
                             "us_cold_l2": round(us, 2), "algorithmic_MB": round(nbytes / 1e6, 2), "GB/s": round(nbytes / (us * 1e-6) / 1e9, 1),
                             "frac": round(nbytes / (us * 1e-6) / 1e9 / hbm, 3)})
    except Exception as ex:                                  # diagnostics only: never lose the headline line over them
        hbm_rows.append({"error": repr(ex)})

    # ---- the same closure on the reference's REAL geometry: fields on a 69 x 721 x 1440 analysis grid over the
    # 128 x 256 network grid (decoder_hr + integrate(interpolation=True), vv_set_case_native), 10 % of the analysis-grid columns
    # observed.  A second number beside the headline, never instead of it; diagnostics only (a failure is recorded, not raised). ----
    native = None
    hbm_used = round((torch.cuda.mem_get_info(dev)[1] - torch.cuda.mem_get_info(dev)[0]) / 2**30, 2)     # of the benchmark workload
    if T > 1 and world == 1:
        try:
            from vaevar_b200.config import era5_stats
            from vaevar_b200.synth import obs_variance
            hr = (721, 1440)
            mean_, std_, _ = era5_stats()
            m_ = torch.from_numpy(mean_).float().to(dev).reshape(1, 69, 1, 1); s_ = torch.from_numpy(std_).float().to(dev).reshape(1, 69, 1, 1)
            gen = torch.Generator(device=dev).manual_seed(0)
            gt_h = m_ + s_ * torch.randn(T, 69, *hr, device=dev, generator=gen)
            xb_h = gt_h[0] + 0.1 * s_[0] * torch.randn(69, *hr, device=dev, generator=gen)
            mask = torch.zeros(hr[0] * hr[1], device=dev)
            mask[torch.randperm(hr[0] * hr[1], device=dev, generator=gen)[: int(args.obs_frac * hr[0] * hr[1])]] = 1.0
            H_h = mask.reshape(1, 1, *hr).expand(T, 69, *hr).contiguous()
            R_h = torch.from_numpy(obs_variance(0.005, 2)).float().to(dev).reshape(1, 69, 1, 1).expand(T, 69, *hr).contiguous()
            eng.set_case_native(xb_h, gt_h, H_h, R_h, 1.0)
            del gt_h, H_h, R_h
            for _ in range(4):
                eng.cost_grad(z, Jb, gb)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                eng.cost_grad(z, Jb, gb)
            e1.record(); torch.cuda.synchronize()
            native = {"analysis_grid": list(hr), "n_obs": int(eng.n_obs), "ms_per_cost_grad": e0.elapsed_time(e1) / 10,
                      "launches": eng.last_launch_count, "J": float(Jb[0]), "grad_finite": bool(torch.isfinite(gb).all())}
            native["over_network_grid"] = native["ms_per_cost_grad"] / ms_eval
        except Exception as ex:
            native = {"error": repr(ex)}
        try:
            eng.set_case(case["xb"], case["yo"], case["H"], case["R"], 1.0)      # back to the benchmark case (and its J)
            eng.cost_grad(z, Jb, gb)
            torch.cuda.synchronize()
        except Exception as ex:
            native = {"error": repr(ex), "restore_failed": True}

    # ---- the cycle's forecast operator at the reference's size: one LGUnet_all_1 application on 69 x 721 x 1440,
    # random-init weights of the shipped architecture; algorithmic flops = its Linears + attention. Diagnostics only. ----
    forecast = None
    if world == 1 and not args.no_forecast:
        try:
            from vaevar_b200.config import FORECAST_FULL
            from vaevar_b200.forecast import ForecastNet
            from vaevar_b200.synth import make_state_dict_net1
            fnet = ForecastNet(FORECAST_FULL, keep_out=69, device=f"cuda:{local}")
            fnet.load_state_dict(make_state_dict_net1(FORECAST_FULL, seed=1)); fnet.finalize()
            xf = torch.randn(69, *FORECAST_FULL.img_size, device=dev)
            for _ in range(2):
                yf = fnet.forward(xf)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                yf = fnet.forward(xf)
            e1.record(); torch.cuda.synchronize()
            ms_f = e0.elapsed_time(e1) / 5
            ops_f = fnet.profile_ops(2)
            fl_f = sum(o["flop"] for o in ops_f) / 1e12
            att = [o for o in ops_f if o["kind"] == "sd_attn" and o["shape"][0] == o["shape"][1]]
            forecast = {"model": "LGUnet_all_1 (networks/LGUnet_all.py:743-777), 69x721x1440, forward only", "ms_per_application": ms_f,
                        "launches": fnet.last_launch_count, "device_GiB": round(fnet.device_bytes / 2**30, 2), "algorithmic_tflop": round(fl_f, 2),
                        "TFLOP/s": round(fl_f / (ms_f * 1e-3), 1), "frac_of_sustained": round(fl_f / (ms_f * 1e-3) / sustained, 3),
                        "whole_grid_attention": {"kernel": "attn1_tc_kernel (tcgen05, TMEM, TMA)", "launches": len(att), "ms": round(sum(o["ms"] for o in att), 3),
                                                 "TFLOP/s": round(sum(o["flop"] for o in att) / max(sum(o["ms"] for o in att), 1e-9) / 1e9, 1)},
                        "finite": bool(torch.isfinite(yf).all())}
            fnet.close()
            del xf, yf
            torch.cuda.empty_cache()
        except Exception as ex:
            forecast = {"error": repr(ex)}

    cycles_per_hour_gpu = (3600.0 / cyc_s) if cyc_s else None
    line = {
        # weak scaling: the per-replica time of one cost+grad (max over ranks); the whole-job aggregate is evals_per_s
        "metric": "ms per 4D-Var cost+grad eval (69x128x256)", "value": ms_eval, "unit": "ms",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_eval, "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp16 forward / bf16 gradients, fp32 accumulate", "data": "synthetic",
        "config": workload_config(T, args.obs_frac),
        "engine": {"recompute": int(args.recompute), "cuda_graph": not args.no_graph, "ln_fold": True, "n_obs": int(eng.n_obs),
                   "parallelism": f"replicas x{world} (independent cases, no data-path collective)"},
        "evals_per_s": 1e3 * args.steps * world / total_ms,
        "cases_in_flight_per_gpu": inflight,
        "da_cycles_per_hour": (cycles_per_hour_gpu * world) if cycles_per_hour_gpu else None,
        "da_cycle": {"measured": bool(cyc_s), "seconds_per_cycle": cyc_s, "closure_evals": cyc_evals,
                     "definition": "one analysis-forecast cycle through vaevar_b200.cycle.CycledDA: Nit=4 x LBFGS.step(max_iter=10, "
                                   "strong Wolfe) + 5 diagnostic sweeps (decode + fused WRMSE/Bias + cost) + 1 forecast step of the flow "
                                   "model on the engine grid (da_4dvar.py:1314-1342, da_4dvar_script.sh:14); identical-twin observations; "
                                   "second of two consecutive cycles, max over ranks, N independent cycle chains in parallel; "
                                   "30-cycle chains: tools/run_cycles.py, profiles/r2_cycles_*.json"},
        "clocks": clocks,
        "e2e": {"value": e2e_ms / args.steps, "unit": "ms", "h2d_bytes_per_step": z_host.numel() * 4,
                "d2h_bytes_per_step": g_host.numel() * 4 + 24,
                "da_cycles_per_hour_per_gpu": cycles_per_hour_gpu, "da_cycles_per_hour_all_gpus": (cycles_per_hour_gpu * world) if cycles_per_hour_gpu else None,
                "seconds_per_da_cycle": cyc_s},
        "gpu_launches": launches * args.steps,
        "gpu_launches_per_step": launches,
        "roofline": {"bound": "tensor", "achieved": gemm_tfs, "peak": sustained, "unit": "TFLOP/s", "frac": gemm_tfs / sustained,
                     "traffic": None,
                     "kernel": "gemm_pair_kernel (tcgen05 cta_group::2, TMEM, TMA) + mlp_fused_kernel (tcgen05 cta_group::1): ALL their launches of the step, time-weighted "
                               "(sum of algorithmic flops / sum of per-launch CUDA-event times, steady state)",
                     "gemm_share_of_step_time": gemm_ms / max(all_ms, 1e-9), "gemm_ms_per_step": gemm_ms,
                     "step_achieved": step_tfs, "step_frac": step_tfs / sustained, "step_frac_of_burst": step_tfs / burst,
                     "peak_source": f"{src} sustained (burst {burst})", "families": roof_rows},
        "roofline_step": {"bound": "tensor", "achieved": step_tfs, "peak": sustained, "unit": "TFLOP/s", "frac": step_tfs / sustained,
                          "algorithmic_tflop": tflop, "peak_source": f"{src} sustained"},
        "roofline_hbm": {"peak": hbm, "unit": "GB/s", "peak_source": src, "l2": "flushed before every timed launch", "kernels": hbm_rows},
        "native_geometry": native,
        "forecast_net": forecast,
        "J": [float(v) for v in Jb.cpu()],
        "hbm_used_gb": hbm_used,
    }
    if not args.no_cpu_baseline and world == 1:
        times, Jcpu, threads = cpu_oracle_eval(T, args.obs_frac, 0, repeats=1)
        line["cpu_baseline"] = {"value": 1e3 * times[0], "unit": "ms", "cores": threads, "kind": "port",
                                "sample": f"1 full closure() (T={T}) of the CPU oracle (reference algorithm, fp32 torch, weight grads off), J={Jcpu:.8g}"}
        try:        # the reference as it ships: weights keep requires_grad=True (da_4dvar.py:590-603), backward fills their gradients too
            t2_, _, _ = cpu_oracle_eval(T, args.obs_frac, 0, repeats=1, as_is=True)
            line["cpu_baseline_as_is"] = {"value": 1e3 * t2_[0], "unit": "ms", "cores": threads, "kind": "port",
                                          "sample": "same closure with weight gradients on, as da_4dvar.py leaves them"}
        except Exception as ex:
            line["cpu_baseline_as_is"] = {"error": repr(ex)}
        # informational: the same PyTorch restatement run eagerly on this B200 (fp32, then TF32 matmuls) -- test infrastructure on the
        # GPU, not the product path
        eager = {}
        if getattr(agent, "_opt", None) is not None:
            agent._opt.close()
        eng.close()
        torch.cuda.empty_cache()
        for name, tf32 in (("fp32", False), ("tf32", True)):
            try:
                te, Je, _ = cpu_oracle_eval(T, args.obs_frac, 0, repeats=3, device=f"cuda:{local}", tf32=tf32)
                eager[name] = {"ms": 1e3 * min(te[1:]), "J": Je}
            except Exception as ex:
                eager[name] = {"error": repr(ex)}
            torch.cuda.empty_cache()
        line["eager_gpu_ms"] = eager
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
