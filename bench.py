#!/usr/bin/env python
"""bench.py -- KBDM solves/s at m = l = 1024 (N = 2048) on 1..8 B200, per the driver contract.

One "step" = one pass of the hot path over ONE FIXED ensemble of `--members` (default 1184 = 8 x 148) members, each the
brain_sim 16-component FID (reference data/params_brain_sim_1_5T.csv, dwell 5e-4) plus its own seeded pseudo-noise draw
(sigma 1e-3), Hankel dimension m = 1024, l = m, p = 1, q = 0 (BASELINE.json: "KBDM solves/sec at L=1024"), solved through the
product's multi-GPU path (llckbdm_b200/distributed.py): longest-processing-time-first shard of the members over the N ranks,
chunked batched solves (llck_kbdm_batched) on every rank, ONE NCCL all_gather of the packed result records.  The ensemble is
fixed, so the 1/2/4/8-GPU curve is STRONG scaling.

  value     whole-job solves/s with the shard's FIDs already resident in HBM: solve + record packing + all_gather on the device
            (CUDA events, max over ranks)
  e2e       the same job through the public host call (distributed.solve_ensemble_distributed: host FIDs in, member-ordered host
            arrays out on every rank; H2D of the shard's FIDs, all_gather, D2H of the gathered records and the re-assembly inside
            the timed region)
  roofline  dominant kernel (hqr_kernel): algorithmic FP64 flops per launch / launch time measured with CUDA events inside the
            timed region (LLCK_FLAG_TIMING) vs the FP64 tensor (DMMA) peak measured on this pool (profiles/fp64_peak_r01.json --
            MEASURED_PEAKS.json carries no FP64 figure)
  weak_148_per_gpu   secondary: every rank solves its own 148-member batch (the round-1 headline), no exchange
  configs   N = 1 only: BASELINE.json configs C1..C5 through the public API (C4 / C5 as scaled subsets, sizes stated), each with
            its fraction of the FP64 peak, and the 148- vs 149-member wave-boundary pair
  cpu_baseline  the reference's CPU path timed on the host cores (baseline/_ref = the unmodified reference when present,
            kind "reference"; else the numpy/scipy restatement in oracle/, kind "port")

`--impl reference` times that CPU path alone (rank 0 only), one m = 1024 solve per step.
"""
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

DWELL = 5e-4
SIGMA = 1e-3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--members", type=int, default=1184, help="members of the fixed ensemble solved per step by the whole job (8 x 148)")
    ap.add_argument("--hankel-dim", dest="m", type=int, default=1024, help="Hankel dimension m = l of every member")
    ap.add_argument("--e2e-steps", type=int, default=3, help="steps of the end-to-end leg (host to host)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1..C5 / wave-boundary extras (N = 1 only)")
    ap.add_argument("--no-weak", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], {}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons[nm] = reasons.get(nm, 0) + 1
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "reason_samples": reasons, "samples": len(sm)}


def fp64_peak_tflops():
    p = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    try:
        d = json.load(open(p))
        return max(d["dmma884_ilp16_tflops"], d["dmma1688_ilp8_tflops"]), "measured on this pool by tools/fp64_peak.cu (DMMA m8n8k4), profiles/fp64_peak_r01.json"
    except Exception:  # noqa: BLE001
        return 37.0, "fallback: nominal B200 FP64 (no measured FP64 peak file)"


def hqr_traffic_per_member(m):
    """dram__bytes_read.sum + dram__bytes_write.sum of one hqr_kernel launch, per member, from the ncu --set full capture of THIS
    round's build (profiles/r02_hqr_traffic.json, written by tools/ncu_summary.py from the capture); None if there is none for m."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_hqr_traffic.json")))
        if int(d["m"]) == int(m):
            return float(d["dram_bytes_per_launch"]) / float(d["members_per_launch"]), d.get("source")
    except Exception:  # noqa: BLE001
        pass
    return None, None


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own code when baseline/_ref holds it (pip install --target of /root/reference, unmodified; the
# numpy >= 1.24 alias it needs is restored by the caller), else the numpy/scipy restatement in oracle/
# ---------------------------------------------------------------------------------------------------------------------
def cpu_solver():
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_dir, "llckbdm")):
        try:
            if not hasattr(np, "complex"):
                np.complex = complex            # noqa: NPY001  (kbdm.py:111-113 uses the alias numpy 1.24 removed)
            sys.path.insert(0, ref_dir)
            from llckbdm.kbdm import kbdm as ref_kbdm
            return (lambda c, m: ref_kbdm(c, DWELL, m=m)), "reference", "the unmodified reference (baseline/_ref, llckbdm.kbdm.kbdm)"
        except Exception:  # noqa: BLE001
            pass
    from oracle.kbdm_oracle import kbdm_oracle
    return (lambda c, m: kbdm_oracle(c, DWELL, m=m, how="einsum")), "port", "numpy/scipy restatement of kbdm.py incl. its 3-operand einsum (oracle/)"


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU reference must run on all host cores, so lift the BLAS limits."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:  # noqa: BLE001
        pass
    return n


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def workload_text(m):
    return f"batched KBDM, brain_sim FID N={2 * m} + pseudo-noise sigma=1e-3 per member, m=l={m}, p=1, q=0"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import scipy.linalg  # noqa: F401  (load BLAS before lifting the thread limits)
    from llckbdm_b200 import workloads
    use_all_host_threads()
    m = args.m
    solve, kind, what = cpu_solver()
    c = workloads.brain_sim(2 * m, SIGMA, 0)
    for _ in range(args.warmup):
        solve(c, m)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solve(c, m)
    dt = time.perf_counter() - t0
    val = args.steps / dt
    cores = blas_threads()
    sample = f"{args.steps} timed solve(s) of one m=l={m} member (N={2 * m}), one per step; {what}"
    print(json.dumps({
        "impl": "reference", "metric": "kbdm_solves_per_sec_m1024", "value": val, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {"workload": workload_text(m), "members_per_step": 1},
        "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_native(args):
    import torch
    import torch.distributed as dist
    from llckbdm_b200 import _native, distributed, ensemble, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; llckbdm_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _native.load()
    m, M = args.m, args.members
    N = 2 * m
    peak, peak_src = fp64_peak_tflops()
    F1 = ensemble.flops_per_solve(m, m)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- the fixed ensemble: one pseudo-noise draw per member, pinned host memory (every rank builds the same job description;
    #      only its own shard is uploaded) ----
    base = workloads.brain_sim(N, 0.0, 0)
    sig_host = torch.empty((M, N), dtype=torch.complex128).pin_memory()
    sig_np = sig_host.numpy()
    for i in range(M):
        g = np.random.default_rng(i)
        sig_np[i] = base + SIGMA * (g.standard_normal(N) + 1j * g.standard_normal(N))
    ms = [m] * M
    plan = distributed.plan_shards(ms, ms)
    mine = plan.mine
    my_dev = sig_host[torch.as_tensor(mine)].to(dev).reshape(-1) if len(mine) else torch.zeros(1, dtype=torch.complex128, device=dev)
    my_off = np.arange(len(mine), dtype=np.int64) * N
    my_len = np.full(len(mine), N, dtype=np.int64)
    torch.cuda.synchronize()

    buf = None
    for _ in range(args.warmup):
        buf = distributed.solve_shard_device(plan, my_dev, my_off, my_len, 1, 0.0, DWELL, buf=buf)
        gathered = distributed.gather_records(plan, buf)
    if args.warmup:
        st = distributed.unpack_records(plan, distributed.records_to_host(gathered))["status"]
        if int((st != 0).sum()) != 0:
            raise RuntimeError("solver reported non-zero status during warm-up")

    # ---- timed region 1: shard FIDs resident in HBM; solve + pack + all_gather ----
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    infos = []
    barrier()
    sampler.start()
    e0.record()
    for _ in range(args.steps):
        buf = distributed.solve_shard_device(plan, my_dev, my_off, my_len, 1, 0.0, DWELL, buf=buf, flags=_native.FLAG_TIMING, infos=infos)
        gathered = distributed.gather_records(plan, buf)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = M * args.steps / (ms_total * 1e-3)
    res = distributed.unpack_records(plan, distributed.records_to_host(gathered))
    bad = int((res["status"] != 0).sum())
    stage_us = np.zeros(9)
    launches, hqr_members = 0, 0
    for nmem, info in infos:
        stage_us += np.array(info[4:13], dtype=float)
        launches += info[13]
        hqr_members += nmem
    launches_all = int(sum_over_ranks(launches))
    n_calls = len(infos)

    # collective alone (same buffers): ms per all_gather
    ag_ms = None
    if world > 1:
        barrier()
        e0.record()
        for _ in range(10):
            distributed.gather_records(plan, buf)
        e1.record()
        barrier()
        ag_ms = max_over_ranks(e0.elapsed_time(e1)) / 10

    # ---- timed region 2: end to end through the public host call ----
    sigs = [sig_np[i] for i in range(M)]
    stats = {}
    ke = max(1, min(args.steps, args.e2e_steps))
    barrier()
    e0.record()
    for _ in range(ke):
        r2 = distributed.solve_ensemble_distributed(sigs, ms, ms, 1, 0.0, DWELL, stats=stats)
    e1.record()
    barrier()
    e2e_val = M * ke / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
    bad += int((r2["status"] != 0).sum())

    # ---- roofline of the dominant kernel ----
    # By device time the top kernel is hqr_kernel (small-bulge multishift QR + AED, one launch per chunk).  Algorithmic flops
    # (SURVEY.md §8d, K5 = 108 l^3 for the whole eig): QR iterations with Schur vectors = 80 l^3 real flops per member (108 l^3
    # minus Hessenberg 13.3 l^3, Q formation 5.3 l^3, eigenvector back-substitution + back-transform 9.3 l^3) -- a NOMINAL
    # (algorithmic) count, not executed DMMA work.  Duration = CUDA events around each launch inside the timed region.
    hqr_s_total = stage_us[5] * 1e-6
    hqr_flops_total = 80.0 * float(m) ** 3 * hqr_members
    achieved = hqr_flops_total / hqr_s_total / 1e12 if hqr_s_total > 0 else 0.0
    per_member_traffic, traffic_src = hqr_traffic_per_member(m)
    members_per_launch = hqr_members / max(1, n_calls)
    gemm_s_total = stage_us[3] * 1e-6
    gemm_flops_total = 16.0 * float(m) ** 3 * hqr_members
    names = ["init_bidiag", "bidiagonal_svd_dc", "backmult", "gemm_T1_Ured", "hessenberg", "hqr", "trevc", "gemm_P_B_W", "epilogue"]

    out = {
        "metric": "kbdm_solves_per_sec_m1024", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {"workload": workload_text(m), "members_per_step": M, "members_per_gpu_per_step": [len(s) for s in plan.shards],
                   "l2": "inputs_larger_than_L2 (per-chunk working set %.1f GB)" % (members_per_launch * 11 * (m * m * 16) / 1e9),
                   "parallelism": f"LPT shard of the fixed ensemble over {world} GPU(s) (llckbdm_b200.distributed), chunked batched solves, "
                                  f"one NCCL all_gather of the packed records per step"},
        "e2e": {"value": e2e_val, "unit": "solves/s", "h2d_bytes_per_step": int(stats.get("h2d_bytes", 0)),
                "d2h_bytes_per_step": int(stats.get("d2h_bytes", 0)), "steps": ke,
                "path": "distributed.solve_ensemble_distributed(host FIDs) -> member-ordered host arrays on every rank"},
        "collective": {"op": "all_gather", "backend": "nccl" if world > 1 else "none (one rank)",
                       "bytes_per_rank": int(plan.count * plan.rec), "bytes_gathered": int(world * plan.count * plan.rec), "ms": ag_ms},
        "gpu_launches": launches_all,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "hqr_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak,
                     "traffic": per_member_traffic * members_per_launch if per_member_traffic else None, "traffic_unit": "B/launch",
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "launches": int(n_calls), "avg_launch_ms": hqr_s_total * 1e3 / max(1, n_calls), "members_per_launch": members_per_launch,
                     "algorithmic_flops_per_launch": hqr_flops_total / max(1, n_calls),
                     "flops_are": "algorithmic (nominal 80 l^3 per member), rank 0's launches"},
        "roofline_secondary": {"bound": "tensor", "kernel": "zgemm_batched_kernel<A_HANKEL> + <A_CONJT> (T1, Ured)",
                               "achieved": gemm_flops_total / gemm_s_total / 1e12 if gemm_s_total > 0 else None, "peak": peak,
                               "unit": "TFLOP/s", "frac": gemm_flops_total / gemm_s_total / 1e12 / peak if gemm_s_total > 0 else None,
                               "launches": int(2 * n_calls)},
        "fp64_roofline_whole_solve": {"algorithmic_tflops_per_gpu": F1 * M * args.steps / (ms_total * 1e-3) / 1e12 / world,
                                      "frac_of_peak_per_gpu": F1 * M * args.steps / (ms_total * 1e-3) / 1e12 / world / peak,
                                      "flops_per_solve": F1},
        "stage_ms_per_step_rank0": {n: float(v) / 1e3 / args.steps for n, v in zip(names, stage_us)},
        "bad_status_members": bad,
    }

    # ---- secondary: weak scaling, every rank its own 148-member batch (the round-1 headline configuration) ----
    if not args.no_weak:
        wb = min(148, M)
        w_dev = my_dev[:wb * N] if len(mine) >= wb else sig_host[:wb].to(dev).reshape(-1)
        w_off = np.arange(wb, dtype=np.int64) * N
        for _ in range(2):
            ensemble.solve_device(w_dev, w_off, [m] * wb, [m] * wb, 1, 0.0, DWELL, sig_len=np.full(wb, N), want_mu=False)
        barrier()
        e0.record()
        for _ in range(3):
            ensemble.solve_device(w_dev, w_off, [m] * wb, [m] * wb, 1, 0.0, DWELL, sig_len=np.full(wb, N), want_mu=False)
        e1.record()
        barrier()
        wms = max_over_ranks(e0.elapsed_time(e1)) / 3
        out["weak_148_per_gpu"] = {"members_per_gpu": wb, "solves_per_s": world * wb / (wms * 1e-3), "ms_per_step": wms,
                                   "frac_of_peak_per_gpu": F1 * wb / (wms * 1e-3) / 1e12 / peak, "scaling": "weak"}

    if rank == 0 and world == 1 and not args.no_configs:
        out["configs"] = run_configs(torch, dev, m, peak)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import scipy.linalg  # noqa: F401
        use_all_host_threads()
        cores = blas_threads()
        solve, kind, what = cpu_solver()
        c = workloads.brain_sim(N, SIGMA, 0)
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            solve(c, m)
            ts.append(time.perf_counter() - t0)
        from oracle.kbdm_oracle import kbdm_oracle
        t0 = time.perf_counter()
        kbdm_oracle(c, DWELL, m=m, how="gemm")
        t_tuned = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1.0 / min(ts), "unit": "solves/s", "cores": cores, "kind": kind,
                               "sample": f"2 solves of one m=l={m} member (N={N}), best of 2; {what}; tuned_value = the oracle port with the "
                                         f"3-operand einsum as GEMM+dot",
                               "tuned_value": 1.0 / t_tuned, "seconds_per_solve": min(ts)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_configs(torch, dev, m, peak):
    """BASELINE.json configs C1..C5 through the public API on ONE GPU (C4 / C5 as scaled subsets) + the wave-boundary pair."""
    from llckbdm_b200 import distributed, ensemble, workloads
    from llckbdm_b200.kbdm import kbdm
    from llckbdm_b200.llckbdm import llc_kbdm
    from llckbdm_b200.min_rmse_kbdm import min_rmse_kbdm
    cfg = {}

    def timed(fn, reps=1):
        best, val = None, None
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            val = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, val

    def frac(ms_, ls_, seconds):
        return float(sum(ensemble.flops_per_solve(a, b) for a, b in zip(ms_, ls_)) / seconds / 1e12 / peak)

    # C1: one kbdm() call, host to host (thread-block clusters per member)
    c1 = workloads.brain_sim(2 * m, SIGMA, 0)
    kbdm(c1, DWELL, m=m)
    t, _ = timed(lambda: kbdm(c1, DWELL, m=m), reps=3)
    cfg["c1_single_solve"] = {"m": m, "seconds": t, "frac_of_peak": frac([m], [m], t), "note": "kbdm(data, dwell, m) host to host, best of 3"}
    # C2: LLC-KBDM ensemble, 100 truncations m in [700, 1024] of one FID: solve phase and the whole llc_kbdm call
    c2 = workloads.brain_sim(2048, SIGMA, 0)
    m2 = workloads.c2_m_range()
    ensemble.solve_ensemble(c2, m2[:4], m2[:4], 1, 0.0, DWELL, device=dev)
    t, r2 = timed(lambda: ensemble.solve_ensemble(c2, m2, m2, 1, 0.0, DWELL, device=dev))
    t_full, r3 = timed(lambda: llc_kbdm(c2, DWELL, m2), reps=2)
    cfg["c2_llc_ensemble"] = {"members": 100, "m_range": "700..1024", "solve_phase_s": t, "solve_frac_of_peak": frac(m2, m2, t),
                              "total_with_clustering_s": t_full, "clusters": int(len(r3.line_list)),
                              "stage_seconds_last_call": dict(__import__("llckbdm_b200.llckbdm", fromlist=["x"]).LAST_STAGE_SECONDS),
                              "bad_status_members": int((r2.status != 0).sum()), "host_cores": os.cpu_count(),
                              "note": "solve_phase_s: host FID in, host line lists out; total_with_clustering_s: llc_kbdm end to end, best of 2 "
                                      "(device solves, pooling, HDBSCAN spanning trees, silhouettes, RMSE selection; dendrogram condensation / "
                                      "EOM labelling on native host threads)"}
    # C3: min_rmse_kbdm sweep over m on a 4096-point FID with a pseudo-noise draw
    c3 = workloads.pseudo_noise_members(workloads.brain_sim(4096, SIGMA, 0), [1])[0]
    m3 = list(range(256, 1025, 64))
    min_rmse_kbdm(c3, DWELL, m_range=m3[:2], l=None)
    t, r = timed(lambda: min_rmse_kbdm(c3, DWELL, m_range=m3, l=None), reps=2)
    cfg["c3_min_rmse_sweep"] = {"members": len(m3), "N": 4096, "m_range": "256..1024 step 64", "seconds": t, "frac_of_peak": frac(m3, m3, t),
                                "min_index": int(r.min_index), "min_rmse": float(r.min_rmse)}
    # C4: MRSI voxels, N = 1024, m = l = 512 -- subset of the 65,536-voxel grid, FIDs synthesised on the device
    nv = 2960
    vox = workloads.c4_voxels_device(0, nv, device=dev).reshape(-1)
    off = np.arange(nv, dtype=np.int64) * 1024
    lens = np.full(nv, 1024, dtype=np.int64)

    def c4():
        bad = 0
        for _idx, r_ in ensemble.solve_chunks(vox, off, lens, [512] * nv, [512] * nv, 1, 0.0, DWELL, want_mu=False):
            bad += int((r_["status"] != 0).sum().item())
            r_["line_lists"].cpu()
        return bad
    t, bad4 = timed(c4)
    cfg["c4_mrsi_voxels"] = {"voxels": nv, "of": 65536, "N": 1024, "m": 512, "seconds": t, "voxels_per_s": nv / t,
                             "frac_of_peak": frac([512] * nv, [512] * nv, t), "bad_status": bad4,
                             "extrapolated_full_grid_s": 65536 * t / nv, "note": "device-resident FIDs, line lists copied to the host"}
    del vox
    # C5: large ragged ensemble through the sharded path (one rank here): m in [512, 1024], own pseudo-noise FID per member
    n5 = 1184
    m5 = workloads.c5_member_sizes(0, n5, stride=37)
    base5 = ensemble.to_device_complex(workloads.brain_sim(4096, SIGMA, 0), dev)
    sig5 = workloads.c5_members_device(base5, 0, n5).cpu().numpy()
    stats = {}
    t, r5 = timed(lambda: distributed.solve_ensemble_distributed([sig5[k] for k in range(n5)], m5, m5, 1, 0.0, DWELL, stats=stats))
    cfg["c5_large_ragged_ensemble"] = {"members": n5, "of": 10000, "m_range": "512..1024", "seconds": t, "members_per_s": n5 / t,
                                       "frac_of_peak": frac(m5, m5, t), "bad_status": int((r5["status"] != 0).sum()),
                                       "extrapolated_10k_members_s": 10000 * t / n5,
                                       "note": "host FIDs in, host records out through distributed.solve_ensemble_distributed"}
    del sig5
    # wave boundary: 148 vs 149 members of m = 1024 (one-CTA-per-member kernels on 148 SMs): as ONE launch sequence, and through the
    # scheduler (ensemble.solve_chunks), which turns the nearly empty second wave into a thread-block-cluster chunk
    sigw = ensemble.to_device_complex(np.concatenate(workloads.pseudo_noise_members(workloads.brain_sim(2 * m, SIGMA, 0), range(149), SIGMA)), dev)
    pair = {}
    for nb in (148, 149):
        offw = np.arange(nb, dtype=np.int64) * 2 * m
        lenw = np.full(nb, 2 * m, dtype=np.int64)

        def one_launch():
            ensemble.solve_device(sigw, offw, [m] * nb, [m] * nb, 1, 0.0, DWELL, sig_len=lenw, want_mu=False)

        def scheduled():
            for _idx, _r in ensemble.solve_chunks(sigw, offw, lenw, [m] * nb, [m] * nb, 1, 0.0, DWELL, want_mu=False):
                pass
        one_launch(); scheduled()
        t1, _ = timed(one_launch, reps=2)
        t2, _ = timed(scheduled, reps=2)
        pair[str(nb)] = {"one_launch_sequence_s": t1, "one_launch_solves_per_s": nb / t1, "scheduler_s": t2, "scheduler_solves_per_s": nb / t2}
    cfg["wave_boundary_m1024"] = pair
    return cfg


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
