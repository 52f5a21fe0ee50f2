"""Stage timing of one batched solve: python tools/time_batch.py m batch [reps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import _native, ensemble
from oracle.kbdm_oracle import brain_sim
m = int(sys.argv[1]); batch = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
N = 2 * m
sigs = [brain_sim(N, 1e-3, seed=i) for i in range(batch)]
flat, offs = ensemble.flatten_signals(sigs, batch)
dev = torch.device("cuda:0")
sig_dev = torch.from_numpy(flat.view(np.float64)).to(dev).view(torch.complex128)
ws = None
names = ["init+bidiag", "jacobi", "final+gather", "T1+Ured", "hessenberg", "hqr", "trevc", "P+B+W", "epilogue"]
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.time()
    out = ensemble.solve_device(sig_dev, offs, [m] * batch, [m] * batch, 1, 0.0, 5e-4, flags=_native.FLAG_TIMING, workspace=ws)
    torch.cuda.synchronize(); dt = time.time() - t0
    ws = out["workspace"]
    info = out["info"]
    st = out["status"].cpu().numpy()
    print(f"m={m} batch={batch} wall={dt:.3f}s solves/s={batch/dt:.2f} jacobi_sweeps={info[0]} max_qr_sweeps={info[1]} bad_status={(st!=0).sum()}")
    print("  " + "  ".join(f"{n}={info[4+i]/1000:.1f}ms" for i, n in enumerate(names)))
    fl = ensemble.flops_per_solve(m, m) * batch
    print(f"  algorithmic TFLOP/s = {fl/dt/1e12:.3f}  ({fl/dt/37.2e12*100:.2f}% of 37.2 TF/s DMMA peak)")
