#!/usr/bin/env python
"""bench.py -- KBDM solves/s at m = l = 1024 (N = 2048) on 1..8 B200, per the driver contract.

One "step" = one pass of the hot path (llck_kbdm_batched) over one batch of synthetic FIDs per GPU:
`--batch` ensemble members (default 148 = one per SM), each the brain_sim 16-component FID (reference
data/params_brain_sim_1_5T.csv, dwell 5e-4) plus its own seeded pseudo-noise draw (sigma 1e-3),
Hankel dimension m = 1024, l = m, p = 1, q = 0  (BASELINE.json: "KBDM solves/sec at L=1024").

  value     whole-job solves/s with the FIDs already resident in HBM (CUDA events, max over ranks)
  e2e       the same through the public host API (ensemble.solve_ensemble: host FIDs in, host line lists out;
            H2D + D2H inside the timed region)
  roofline  dominant kernel (hqr_kernel): algorithmic FP64 flops per launch / measured launch time
            vs the FP64 tensor (DMMA) peak measured on this pool (profiles/fp64_peak_r01.json -- MEASURED_PEAKS.json
            carries no FP64 figure)
  llc_ensemble_c2  config C2 (100 truncations m in [700,1024]): solve-phase and end-to-end llc_kbdm latency
  single_solve_c1  config C1: one kbdm() call, host to host
  cpu_baseline  the numpy/scipy restatement of the reference (oracle/, kind "port") timed on the host cores

`--impl reference` times that CPU restatement alone (rank 0 only), one m = 1024 solve per step.
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
    ap.add_argument("--batch", type=int, default=148, help="ensemble members per GPU per step")
    ap.add_argument("--m", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c2", action="store_true")
    ap.add_argument("--no-c2-full", action="store_true", help="skip the end-to-end llc_kbdm timing of config C2 (solve + clustering, ~10 s)")
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
        sm, mx, reasons = [], [], set()
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
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fp64_peak_tflops():
    p = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    try:
        d = json.load(open(p))
        return max(d["dmma884_ilp16_tflops"], d["dmma1688_ilp8_tflops"]), "measured on this pool by tools/fp64_peak.cu (DMMA m8n8k4), profiles/fp64_peak_r01.json"
    except Exception:  # noqa: BLE001
        return 37.0, "fallback: nominal B200 FP64 (no measured FP64 peak file)"


def cpu_reference_solve(m, how):
    """One solve of the reference's CPU algorithm (oracle port) at Hankel size m on all host cores."""
    from oracle.kbdm_oracle import brain_sim, kbdm_oracle
    c = brain_sim(2 * m, SIGMA, 0)
    t0 = time.perf_counter()
    kbdm_oracle(c, DWELL, m=m, how=how)
    return time.perf_counter() - t0


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import scipy.linalg  # noqa: F401  (load BLAS before lifting the thread limits)
    use_all_host_threads()
    m = args.m
    for _ in range(args.warmup):
        cpu_reference_solve(m, "einsum")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_solve(m, "einsum")
    dt = time.perf_counter() - t0
    val = args.steps / dt
    cores = blas_threads()
    sample = f"{args.steps} timed solve(s) of one m=l={m} member (N={2 * m}) per step; numpy/scipy restatement of kbdm.py incl. its 3-operand einsum"
    print(json.dumps({
        "impl": "reference", "metric": "kbdm_solves_per_sec_m1024", "value": val, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {"workload": f"batched KBDM, brain_sim FID N={2 * m} + pseudo-noise sigma=1e-3, m=l={m}, p=1, q=0", "members_per_step": 1},
        "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_native(args):
    import torch
    import torch.distributed as dist
    from llckbdm_b200 import _native, ensemble
    from oracle.kbdm_oracle import brain_sim      # input generator + cpu_baseline leg only

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
    m, batch = args.m, args.batch
    N = 2 * m

    # ---- synthetic inputs: one pseudo-noise draw per member, pinned host memory ----
    sig_host = torch.empty((batch, N), dtype=torch.complex128).pin_memory()
    sig_np = sig_host.numpy()
    for i in range(batch):
        sig_np[i] = brain_sim(N, SIGMA, seed=rank * batch + i)
    offsets = np.arange(batch, dtype=np.int64) * N
    ms = [m] * batch
    sig_dev = sig_host.to(dev, non_blocking=True).reshape(-1)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ws = None
    last = None
    for _ in range(args.warmup):
        last = ensemble.solve_device(sig_dev, offsets, ms, ms, 1, 0.0, DWELL, workspace=ws)
        ws = last["workspace"]
    if last is not None and int((last["status"] != 0).sum().item()) != 0:
        raise RuntimeError("solver reported non-zero status during warm-up")

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record()
    stage_us = np.zeros(9)
    launches = 0
    for _ in range(args.steps):
        last = ensemble.solve_device(sig_dev, offsets, ms, ms, 1, 0.0, DWELL, workspace=ws, flags=_native.FLAG_TIMING)
        info = last["info"]
        stage_us += np.array(info[4:13], dtype=float)
        launches += info[13]
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * batch * args.steps / (ms_total * 1e-3)
    bad = int((last["status"] != 0).sum().item())

    # ---- timed region 2: end to end through the public host API ----
    h2d = batch * N * 16
    d2h = batch * (m * 4 * 8 + 2 * m * 16 + m * 8 + 8)
    barrier()
    e0.record()
    for _ in range(args.steps):
        res = ensemble.solve_ensemble([sig_np[i] for i in range(batch)], ms, ms, 1, 0.0, DWELL, device=dev, chunk=batch)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * batch * args.steps / (float(t.item()) * 1e-3)
    bad += int((res.status != 0).sum())

    # ---- roofline of the dominant kernel ----
    # By device time the top kernel is hqr_kernel (small-bulge multishift QR + AED, one launch per step, ~24 % of the step).
    # Algorithmic flops (SURVEY.md §8d, K5 = 108 l^3 for the whole eig): QR iterations with Schur vectors = 80 l^3 real flops per
    # member (108 l^3 minus Hessenberg 13.3 l^3, Q formation 5.3 l^3, eigenvector back-substitution + back-transform 9.3 l^3).
    # Duration = CUDA events around the launch inside the timed region (stage timer info[9]).
    peak, peak_src = fp64_peak_tflops()
    hqr_s = (stage_us[5] * 1e-6) / args.steps
    flops_per_launch = 80.0 * float(m) ** 3 * batch
    achieved = flops_per_launch / hqr_s / 1e12
    # secondary: the two large DMMA GEMMs of the reduced operator (T1 = U^p Rs with the implicit-Hankel A operand, Ured = Lt^H T1):
    # 16 m^3 real flops per member, timed by the stage events around the two launches
    gemm_flops = 16.0 * float(m) ** 3 * batch
    gemm_s = (stage_us[3] * 1e-6) / args.steps
    alg_flops = ensemble.flops_per_solve(m, m) * batch * args.steps
    names = ["init_bidiag", "bidiagonal_svd_dc", "finalize_backmult", "gemm_T1_Ured", "hessenberg", "hqr", "trevc", "gemm_P_B_W", "epilogue"]

    out = {
        "metric": "kbdm_solves_per_sec_m1024", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {"workload": f"batched KBDM (config: LLC ensemble members as pseudo-noise draws), brain_sim FID N={N} sigma=1e-3, m=l={m}, p=1, q=0",
                   "members_per_gpu_per_step": batch, "l2": "inputs_larger_than_L2 (per-step working set %.1f GB)" % (batch * 11 * (m * m * 16) / 1e9),
                   "parallelism": f"members sharded over {world} GPU(s), no data-path collective"},
        "e2e": {"value": e2e_val, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "hqr_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch (148 members, m=1024) from the ncu --set full capture
                     # summarised in profiles/r01t_ncu_summary.md (561.3 GB read + 547.1 GB written); None for other shapes
                     "traffic": 1.1084e12 if (batch == 148 and m == 1024) else None, "traffic_unit": "B/launch", "peak_source": peak_src,
                     "launches": int(args.steps), "avg_launch_ms": hqr_s * 1e3,
                     "algorithmic_flops_per_launch": flops_per_launch},
        "roofline_secondary": {"bound": "tensor", "kernel": "zgemm_batched_kernel<A_HANKEL> + <A_CONJT> (T1, Ured)",
                               "achieved": gemm_flops / gemm_s / 1e12, "peak": peak, "unit": "TFLOP/s",
                               "frac": gemm_flops / gemm_s / 1e12 / peak, "launches": int(2 * args.steps),
                               "avg_launch_ms": gemm_s * 1e3 / 2, "algorithmic_flops_per_launch": gemm_flops / 2},
        "fp64_roofline_whole_solve": {"algorithmic_tflops": alg_flops / (ms_total * 1e-3) / 1e12 / 1.0,
                                      "frac_of_peak_per_gpu": alg_flops / (ms_total * 1e-3) / 1e12 / peak,
                                      "flops_per_solve": ensemble.flops_per_solve(m, m)},
        "stage_ms_per_step": {n: float(v) / 1e3 / args.steps for n, v in zip(names, stage_us)},
        "bad_status_members": bad,
    }

    if rank == 0 and world == 1 and not args.no_c2:
        # LLC-KBDM ensemble (config C2): 100 truncations m in [700,1024] of one FID, solve-phase latency (clustering stays on CPU)
        c = brain_sim(2048, SIGMA, 0)
        m2 = [700 + round(k * 324 / 99) for k in range(100)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r2 = ensemble.solve_ensemble(c, m2, m2, 1, 0.0, DWELL, device=dev)
        torch.cuda.synchronize()
        out["llc_ensemble_c2"] = {"members": 100, "m_range": "700..1024", "solve_phase_s": time.perf_counter() - t0,
                                  "bad_status_members": int((r2.status != 0).sum()), "note": "host FID in, host line lists out; HDBSCAN clustering not included"}
    if rank == 0 and world == 1 and not args.no_c2:
        # single KBDM solve (config C1): one m = l = 1024 member through the public API, host to host (thread-block clusters per member)
        from llckbdm_b200.kbdm import kbdm
        c1 = brain_sim(2 * m, SIGMA, 0)
        kbdm(c1, DWELL, m=m)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            kbdm(c1, DWELL, m=m)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        out["single_solve_c1"] = {"m": m, "seconds": float(np.median(ts)), "note": "kbdm(data, dwell, m) host to host, median of 3"}
    if rank == 0 and world == 1 and not args.no_c2 and not args.no_c2_full:
        from llckbdm_b200.llckbdm import llc_kbdm
        c = brain_sim(2048, SIGMA, 0)
        m2 = [700 + round(k * 324 / 99) for k in range(100)]
        t0 = time.perf_counter()
        r3 = llc_kbdm(c, DWELL, m2)
        out["llc_ensemble_c2"]["total_with_clustering_s"] = time.perf_counter() - t0
        out["llc_ensemble_c2"]["note"] = ("solve_phase_s: host FID in, host line lists out; total_with_clustering_s: llc_kbdm end to end "
                                          "(device solves, HDBSCAN spanning trees, silhouettes and RMSE selection; tree condensation on the host cores)")
        out["llc_ensemble_c2"]["clusters"] = int(len(r3.line_list))
        out["llc_ensemble_c2"]["host_cores"] = os.cpu_count()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import scipy.linalg  # noqa: F401
        use_all_host_threads()
        cores = blas_threads()
        t_ref = cpu_reference_solve(m, "einsum")
        t_tuned = cpu_reference_solve(m, "gemm")
        out["cpu_baseline"] = {"value": 1.0 / t_ref, "unit": "solves/s", "cores": cores, "kind": "port",
                               "sample": f"1 solve of one m=l={m} member (N={N}), numpy/scipy restatement of kbdm.py as written (3-operand einsum); "
                                         f"tuned_value = same with the einsum as GEMM+dot",
                               "tuned_value": 1.0 / t_tuned, "seconds_per_solve": t_ref}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
