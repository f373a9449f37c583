#!/usr/bin/env python
"""bench.py — suggestions/sec at (N=8192 trials, M=1,048,576 candidates, D=32), Matérn-5/2 GP, EI  (BASELINE.json cfg 3).

One "step" = one full suggestion at fixed θ: Gram → Cholesky → L⁻¹/alpha → sweep of this rank's candidate rows
(K*, mean, tcgen05 variance contraction) → EI → first-index argmax (→ NCCL argmax all-reduce when N > 1).

  value     device-resident inputs, CUDA-event timed, max over ranks.
  e2e       the same through kbo_suggest_host: HOST buffers in (pinned), H2D + 32-byte D2H inside the timed region.
  roofline  tcgen05 variance kernel: algorithmic flops (rows·N² per launch) ÷ its mean launch time (CUDA events
            recorded around every launch inside libkbo) against the measured bf16 tensor peak; plus the standalone
            acquisition pass against the measured HBM copy bandwidth ("acquisition HBM GB/s vs peak").
  cpu_baseline / --impl reference   the oracle (fp64 NumPy/SciPy port of the sklearn path; scikit-optimize itself is
            not installable here) on the host cores, bounded sample, extrapolated linearly in M (the sweep is).

Multi-GPU (torchrun, one rank per GPU): weak scaling — every rank runs the full N=8192 fit (replicated,
bit-identical) and sweeps its own 1,048,576-row shard of an R·1,048,576 grid; `value` counts R
(N=8192, M=1M)-suggestion units per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRIALS, M_CAND, DIM = 8192, 1_048_576, 32
KERNEL, ACQ = "matern52", "ei"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


def tc_traffic(rows_per_launch):
    """dram bytes per launch of tc_variance_kernel from the committed `ncu --set full` capture (profiles/), scaled by rows."""
    p = os.path.join(ROOT, "profiles", "tcvar_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    out = {"bytes_per_launch": (d["dram_read_bytes"] + d["dram_write_bytes"]) * rows_per_launch / d["rows"], "source": d["source"],
           "captured_rows_per_launch": d["rows"]}
    for k in ("algorithmic_bytes_per_launch", "ncu_tensor_pipe_active_pct_of_active", "ncu_dram_pct_of_peak", "ncu_l2_hit_pct",
              "why_traffic_exceeds_algorithmic"):
        if k in d:
            out[k] = d[k]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_baseline(sample_rows: int = 2048, do_fit: bool = True):
    """Oracle on the host cores: full N=8192 fit + a `sample_rows`-candidate sweep, extrapolated linearly to M."""
    from oracle import gp_oracle as O
    try:
        from threadpoolctl import threadpool_info
        threads = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        threads = os.cpu_count() or 1
    X, y, Xc = O.synthetic(N_TRIALS, sample_rows, DIM)
    th = O.theta_of_record(DIM)
    t0 = time.perf_counter()
    fit = O.gp_fit(X, y, kind=KERNEL, length_scale=th["length_scale"], amplitude=th["amplitude"], noise=th["noise"])
    t_fit = time.perf_counter() - t0
    t0 = time.perf_counter()
    mu, std = O.gp_predict(fit, Xc, batch=sample_rows)
    a = O.acquisition(mu, std, float(y.min()), ACQ, th["xi"], th["kappa"])
    _ = O.first_argmax(a)
    t_sweep = time.perf_counter() - t0
    t_full = t_fit + t_sweep * (M_CAND / sample_rows)
    return {"value": 1.0 / t_full, "unit": "suggestions/s", "cores": int(threads), "kind": "port",
            "sample": f"full fit N={N_TRIALS} ({t_fit:.2f}s) + {sample_rows}-candidate sweep ({t_sweep:.2f}s) extrapolated linearly to M={M_CAND}",
            "fit_s": t_fit, "sweep_sample_s": t_sweep, "host_cpus": os.cpu_count()}


def run_reference(args, rank):
    if rank != 0:
        return
    # each step: the bounded sample (fit once in warm-up is NOT reused: every step pays the full fit like the reference does)
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(sample_rows=2048)
        if i >= args.warmup:
            vals.append(1.0 / last["value"])
    t = float(np.mean(vals))
    line = {"metric": "suggestions/sec at (N=8192, M=1M, D=32)", "value": 1.0 / t, "unit": "suggestions/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg3: GP(Matern-5/2) N=8192 D=32, EI sweep over M=1048576 candidates, fixed theta", "kernel": KERNEL, "acq": ACQ},
            "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": 1.0 / t, "unit": "suggestions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["cpu_baseline"]["value"] = line["value"]
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trials", type=int, default=N_TRIALS)       # for quick local experiments only; the bench of record uses defaults
    ap.add_argument("--candidates", type=int, default=M_CAND)
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--var-mode", default="tc", choices=["tc", "f64"])
    ap.add_argument("--k-span", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from kubeflow_b200.gp import GPEngine
    from kubeflow_b200.dist import global_argmax
    from oracle import gp_oracle as O   # only for the synthetic workload definition and the cpu_baseline leg

    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's version banner off stdout: stdout carries ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, M, D = args.trials, args.candidates, args.dim
    th = O.theta_of_record(D)
    X, y, _ = O.synthetic(N, 1, D)
    # this rank's shard of the R·M grid: rows [rank·M, (rank+1)·M) of rng(4321).random((R·M, D)), fp32 candidates
    r = np.random.default_rng(4321)
    if rank:
        r.random((rank * M, D))   # advance the stream to this rank's rows (same values as slicing the full grid)
    Xc = r.random((M, D)).astype(np.float32)
    goff = rank * M

    eng = GPEngine(local, kernel=KERNEL, acq=ACQ, var_mode=args.var_mode, tc_k_span=args.k_span, **th)
    dev = torch.device("cuda", local)
    Xd, yd, Xcd = torch.tensor(X, device=dev), torch.tensor(y, device=dev), torch.tensor(Xc, device=dev)

    def step_device():
        eng.tell(Xd, yd)
        b = eng.ask(Xcd, global_offset=goff)
        return global_argmax(b) if world > 1 else b

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident value ---------------------------------------------------------------------------------
    for _ in range(args.warmup):
        best = step_device()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        best = step_device()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / args.steps

    # ---- end to end: host buffers through kbo_suggest_host ---------------------------------------------------------
    Xp, yp = torch.tensor(X).pin_memory(), torch.tensor(y).pin_memory()
    Xcp = torch.tensor(Xc).pin_memory()
    Xh, yh, Xch = Xp.numpy(), yp.numpy(), Xcp.numpy()

    def step_host():
        b, t = eng.suggest_host(Xh, yh, Xch, global_offset=goff)
        return (global_argmax(b) if world > 1 else b), t

    for _ in range(args.warmup):
        best_h, tim = step_host()
    barrier()
    e0.record()
    var_ms, cross_ms, acq_ms_list, fit_ms, cal_ms, launches, chunks = [], [], [], [], [], 0, 0
    for _ in range(args.steps):
        best_h, tim = step_host()
        var_ms.append(tim["var_kernel_ms"]); cross_ms.append(tim["cross_kernel_ms"]); acq_ms_list.append(tim["acq_kernel_ms"])
        fit_ms.append(tim["fit_ms"]); cal_ms.append(tim["calib_ms"]); launches += tim["launches"]; chunks = tim["chunks"]
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms_step_e2e = float(ms2.item()) / args.steps
    assert best_h.index == best.index, "host and device paths disagree on the argmax"

    # ---- standalone acquisition pass: the HBM-bound kernel (8 B/candidate in + 4 B out), at this rank's M and at cfg5's 16M ----
    flush64 = torch.zeros((256 << 20) // 8, dtype=torch.int64, device=dev)   # > 126 MB L2
    flush_sink = torch.zeros((), dtype=torch.int64, device=dev)

    def acq_bw(Ma):
        mu_n = torch.randn(Ma, device=dev, dtype=torch.float32)
        var_n = torch.rand(Ma, device=dev, dtype=torch.float32)
        acq_o = torch.empty(Ma, device=dev, dtype=torch.float32)
        ts = []
        for i in range(10):
            # evict with READS of a 256 MiB buffer: a write-based flush leaves ~126 MB of dirty L2 lines whose write-back then
            # competes with the timed kernel for DRAM bandwidth (measured: 3.3 -> TB/s reading of the same kernel)
            flush_sink.copy_(flush64.sum())
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            eng.lib.kbo_acq_argmax(eng._h, mu_n.data_ptr(), var_n.data_ptr(), Ma, 0, 0, 0.0, 1.0, -1.0, 0.01, 1.96, acq_o.data_ptr(),
                                   eng._best_dev.data_ptr(), eng._stream())
            a1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a0.elapsed_time(a1))
        t = float(np.median(ts))
        return 12.0 * Ma / (t * 1e-3) / 1e9, t

    acq_gbs, acq_ms = acq_bw(M)
    acq_gbs16, acq_ms16 = acq_bw(16_777_216)

    # ---- the other BASELINE.json configurations, briefly (rank 0, single GPU): cfg2 (GP RBF) and cfg4 (CMA-ES) ---------------------
    other = {}
    if rank == 0 and world == 1:
        try:
            X2, y2, Xc2 = O.synthetic(1024, 65536, 8)
            e2 = GPEngine(local, kernel="rbf", acq="ei", var_mode="auto", **O.theta_of_record(8))
            X2d, y2d, Xc2d = torch.tensor(X2, device=dev), torch.tensor(y2, device=dev), torch.tensor(Xc2.astype(np.float32), device=dev)
            for _ in range(3):
                e2.tell(X2d, y2d); e2.ask(Xc2d)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                e2.tell(X2d, y2d); b2 = e2.ask(Xc2d)
            torch.cuda.synchronize()
            other["cfg2_gp_rbf_n1024_m65536_d8"] = {"suggestions_per_s": 10 / (time.perf_counter() - t0), "var_mode": "auto (FP64 contraction)",
                                                     "argmax_index": b2.index}
            e2.close()
            from kubeflow_b200.cmaes import CmaEs
            es = CmaEs(np.full(128, 3.0), 2.0, popsize=4096, seed=7, device=local)
            es.run_synthetic("rastrigin", 20)
            r4 = es.run_synthetic("rastrigin", 200)
            other["cfg4_cmaes_d128_pop4096_200gen"] = {"generations_per_s": r4["generations_per_s"], "elapsed_ms": r4["elapsed_ms"],
                                                        "jacobi_sweeps_last": r4["jacobi_sweeps"]}
            es.close()
            from oracle import cma_oracle as CO
            st = CO.CmaState(np.full(128, 3.0), 2.0, 4096)
            rr = np.random.default_rng(7)
            t0 = time.perf_counter()
            for _ in range(5):
                Xo, Yo = CO.ask(st, rr.standard_normal((4096, 128)))
                CO.tell(st, Yo, CO.rastrigin(Xo))
            other["cfg4_cmaes_d128_pop4096_200gen"]["cpu_oracle_generations_per_s"] = 5 / (time.perf_counter() - t0)
            # the reference-facing surface at cfg3 scale: a real gRPC GetSuggestions carrying all 8192 finished trials as strings
            import grpc
            from kubeflow_b200.suggestion import api_pb as api
            from kubeflow_b200.suggestion.server import SuggestionStub, serve
            from kubeflow_b200.suggestion.service import SkoptService
            skopt_svc = SkoptService({"device": local})
            server, port = serve(skopt_svc, port=0, host="127.0.0.1")
            ch = grpc.insecure_channel(f"127.0.0.1:{port}", options=[("grpc.max_send_message_length", 1 << 28), ("grpc.max_receive_message_length", 1 << 28)])
            stub = SuggestionStub(ch)
            ex = api.Experiment()
            ex.name = "bench-cfg3"
            ex.spec.objective.type = api.MINIMIZE
            ex.spec.objective.objective_metric_name = "loss"
            ex.spec.algorithm.algorithm_name = "bayesianoptimization"
            for k_, v_ in {"n_initial_points": 0, "acq_func": "EI", "acq_optimizer": "sampling", "random_state": 1, "n_points": M, "var_mode": "tc"}.items():
                st_ = ex.spec.algorithm.algorithm_settings.add()
                st_.name, st_.value = k_, str(v_)
            for d_ in range(D):
                ps_ = ex.spec.parameter_specs.parameters.add()
                ps_.name, ps_.parameter_type = f"x{d_}", api.DOUBLE
                ps_.feasible_space.min, ps_.feasible_space.max = "0", "1"
            rq = api.GetSuggestionsRequest(experiment=ex, current_request_number=1)

            def add_trial(i, xs, yv):
                t_ = rq.trials.add()
                t_.name = f"t{i}"
                t_.spec.objective.objective_metric_name = "loss"
                t_.status.condition = api.SUCCEEDED
                for d_ in range(D):
                    a_ = t_.spec.parameter_assignments.assignments.add()
                    a_.name, a_.value = f"x{d_}", repr(float(xs[d_]))
                m_ = t_.status.observation.metrics.add()
                m_.name, m_.value = "loss", repr(float(yv))

            for i in range(N - 3):
                add_trial(i, X[i], y[i])
            calls = []
            for c_ in range(4):
                t0 = time.perf_counter()
                rep = stub.GetSuggestions(rq)
                calls.append((time.perf_counter() - t0) * 1e3)
                add_trial(N - 3 + c_, X[min(N - 3 + c_, N - 1)], y[min(N - 3 + c_, N - 1)])
            other["grpc_cfg3_request"] = {"trials_in_request": N, "request_bytes": rq.ByteSize(), "n_points": M, "cold_call_ms": calls[0],
                                          "steady_call_ms": float(np.median(calls[1:])),
                                          "ingest_last_call": skopt_svc.last_ingest,
                                          "engine_update_last_call": getattr(skopt_svc._services["bench-cfg3"].skopt_optimizer, "last_fit", None),
                                          "note": "in-process grpc.server, all finished trials resent as strings on every call; steady = one new trial per call"}
            ch.close()
            server.stop(0)
        except Exception as e:  # noqa: BLE001 — the extras must never take the headline line down
            other["error"] = f"{type(e).__name__}: {e}"

    if rank == 0:
        pk = peaks()
        rows_per_launch = M / max(chunks, 1)
        var_launch_ms = float(np.mean(var_ms)) / max(chunks, 1)
        flops_launch = rows_per_launch * float(N) * float(N)           # Σ_j Σ_{k<=j} 2 flops = N² per candidate row
        achieved_tf = flops_launch / (var_launch_ms * 1e-3) / 1e12 if args.var_mode == "tc" else flops_launch / (var_launch_ms * 1e-3) / 1e12
        rank_err = eng.last_rank_error() if args.var_mode == "tc" else 0.0
        fast_rank = args.var_mode == "tc" and rank_err > 0.0
        mma_products = 1.0 if fast_rank else 3.0
        line = {
            "metric": "suggestions/sec at (N=8192, M=1M, D=32)", "value": world * 1e3 / ms_step, "unit": "suggestions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": ("f64 fit+mean / fp16 tcgen05 ranking pass over sigma² (1 product, fp32 accumulate) + FP64 decision among the survivors"
                      if fast_rank else "f64 fit+mean / fp16x3-split tcgen05 variance (fp32 accumulate)") if args.var_mode == "tc" else "f64",
            "data": "synthetic",
            "config": {"workload": f"cfg3: GP({KERNEL}) N={N} D={D}, {ACQ.upper()} sweep over M={M} candidates per GPU (grid {world * M}), fixed theta "
                                   f"(amp 1, ls 0.3*sqrt(D), noise 1e-3), one suggestion per step",
                       "kernel": KERNEL, "acq": ACQ, "per_gpu_candidates": M, "grid_candidates": world * M, "var_mode": args.var_mode,
                       "l2": "inputs larger than L2 (Xc 128 MiB, W planes 256 MiB, K* scratch ~2 GiB per chunk)", "argmax_index": best.index,
                       "fp64_refined_contenders": eng.last_contenders() if args.var_mode == "tc" else None,
                       "ranking_pass": ("1 fp16 product per term, error bound calibrated per sweep: max |d sigma²| on the calibration rows = %.3g" % rank_err) if fast_rank else None},
            "e2e": {"value": world * 1e3 / ms_step_e2e, "unit": "suggestions/s", "ms_per_step": ms_step_e2e,
                    "h2d_bytes_per_step": int(X.nbytes + y.nbytes + Xc.nbytes), "d2h_bytes_per_step": 32},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "tc_variance_pair_kernel" if os.environ.get("KBO_TC_PAIR", "1") != "0" else "tc_variance_kernel", "achieved": achieved_tf, "peak": pk["tf_sus"], "unit": "TFLOP/s",
                         "frac": achieved_tf / pk["tf_sus"], "peak_source": f"bf16_tflops_sustained, {pk['src']}",
                         "issued_mma_tflops": mma_products * achieved_tf * (1.0 + 256.0 / N), "mma_products_per_term": mma_products,
                         "traffic": tc_traffic(rows_per_launch) if not fast_rank else None,
                         "traffic_note": None if not fast_rank else "the ncu --set full capture on file (profiles/tcvar_traffic.json: 31.8 GB per launch) is of the three-product kernel; the ranking pass loads the hi planes only, i.e. half of that",
                         "launch_ms": var_launch_ms, "launches_per_step": chunks, "flops_per_launch": flops_launch,
                         "launch_note": ("variance phase time / chunks; the phase also holds the calibration launch (first wave of rows, three products), "
                                         "so the per-launch time is overstated by ~5 %") if fast_rank else None},
            "acquisition_hbm": {"kernel": "acq_kernel<float>", "bytes_per_candidate": 12, "candidates": M, "achieved": acq_gbs, "peak": pk["hbm"],
                                "unit": "GB/s", "frac": acq_gbs / pk["hbm"], "peak_source": f"hbm_gbs, {pk['src']}", "launch_ms": acq_ms,
                                "at_16M_candidates": {"achieved": acq_gbs16, "frac": acq_gbs16 / pk["hbm"], "launch_ms": acq_ms16},
                                "l2": "evicted by reading a 256 MiB buffer before each timed launch (clean lines: no write-back during the timed kernel)"},
            "other_configs": other,
            "phases_ms": {"fit": float(np.mean(fit_ms)), "cross_kernel": float(np.mean(cross_ms)), "variance_kernel": float(np.mean(var_ms)),
                          "calibration": float(np.mean(cal_ms)), "acquisition_and_fp64_decision": float(np.mean(acq_ms_list))},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
