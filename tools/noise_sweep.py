"""Parity vs the oracle as the FID noise level (hence cond(U0)) varies: D&C SVD (default) against the Jacobi back end.
    python tools/noise_sweep.py [m]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from llckbdm_b200 import _native
from llckbdm_b200.ensemble import solve_ensemble
from oracle.kbdm_oracle import brain_sim, kbdm_oracle, compare_members
m = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for sigma in (1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-9):
    c = brain_sim(2 * m, sigma, 0)
    _, info, mu, D = kbdm_oracle(c, 5e-4, m=m, return_mu=True)
    s = info.singular_values
    row = [f"sigma={sigma:.0e} smin/smax={s[-1] / s[0]:.1e}"]
    for mode, opt in (("dc", _native.SVD_DC), ("jacobi", _native.SVD_JACOBI)):
        res = solve_ensemble(c, [m], [m], 1, 0.0, 5e-4, options=_native.Options(svd_mode=opt))
        dmu, dD = compare_members(res.mu[0, :m], res.D[0, :m], mu, D)
        dsv = np.max(np.abs(res.sing_vals[0, :m] - s) / s)
        row.append(f"{mode}: status={int(res.status[0])} dmu={dmu:.1e} dD={dD:.1e} dsv={dsv:.1e}")
    print("  ".join(row), flush=True)
