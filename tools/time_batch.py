"""Stage timing of one batched solve (+ the clock64 phase split of hqr_kernel): python tools/time_batch.py m batch [reps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import _native, ensemble
from oracle.kbdm_oracle import brain_sim
m = int(sys.argv[1]); batch = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
N = 2 * m
sigs = [brain_sim(N, 1e-3, seed=i) for i in range(batch)]
flat, offs, lens = ensemble.flatten_signals(sigs, batch)
dev = torch.device("cuda:0")
sig_dev = torch.from_numpy(flat.view(np.float64)).to(dev).view(torch.complex128)
names = ["init+bidiag", "svd_bidiagonal", "backmult", "T1+Ured", "hessenberg", "hqr", "trevc", "P+B+W", "epilogue"]
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.time()
    prof = torch.zeros((batch, 10), dtype=torch.int64, device=dev)
    opts = _native.Options(hqr_profile=prof.data_ptr())
    out = ensemble.solve_device(sig_dev, offs, [m] * batch, [m] * batch, 1, 0.0, 5e-4, flags=_native.FLAG_TIMING, sig_len=lens,
                                options=opts)
    torch.cuda.synchronize(); dt = time.time() - t0
    info = out["info"]
    st = out["status"].cpu().numpy()
    print(f"m={m} batch={batch} wall={dt:.3f}s solves/s={batch/dt:.2f} launches={info[13]} max_qr_sweeps={info[1]} bad_status={(st!=0).sum()}")
    pm = prof.cpu().numpy().astype(float).mean(axis=0) / 1e6
    print("  hqr Mcycles/member: aed_schur=%.1f load=%.1f chase=%.1f store=%.1f strips=%.1f small=%.1f aed_reorder=%.1f aed_warp=%.1f aed_strips=%.1f aed_calls=%.0f"
          % (pm[0], pm[1], pm[2], pm[3], pm[4], pm[5], pm[6], pm[7], pm[8], pm[9] * 1e6))
    print("  " + "  ".join(f"{n}={info[4+i]/1000:.1f}ms" for i, n in enumerate(names)))
    fl = ensemble.flops_per_solve(m, m) * batch
    print(f"  algorithmic TFLOP/s = {fl/dt/1e12:.3f}  ({fl/dt/37.2e12*100:.2f}% of 37.2 TF/s DMMA peak)")
