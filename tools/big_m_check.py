import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from llckbdm_b200.ensemble import solve_ensemble
from llckbdm_b200.kbdm import kbdm
from oracle.kbdm_oracle import brain_sim, kbdm_oracle, compare_members
c = brain_sim(4096, 1e-3, 0)
for m in (1536, 2048):
    t0 = time.time(); res = solve_ensemble(c, [m, m - 7], [m, m - 7], 1, 0.0, 5e-4); dt = time.time() - t0
    print("m", m, "status", res.status, "time", round(dt, 2), "sv0", res.sing_vals[0, :2], flush=True)
    if m == 1536:
        _, info, mu, D = kbdm_oracle(c, 5e-4, m=m, return_mu=True)
        print("parity m=1536:", compare_members(res.mu[0, :m], res.D[0, :m], mu, D), np.max(np.abs(res.sing_vals[0, :m] - info.singular_values) / info.singular_values))
# NaN in the unused tail must not matter; NaN in the used part raises before launch
c2 = brain_sim(512, 1e-3, 0); c2[400] = np.nan
ll, _ = kbdm(c2, 5e-4, m=100); print("tail NaN ok", np.isfinite(ll).all())
try:
    kbdm(c2, 5e-4, m=250)
except ValueError as e:
    print("ValueError:", e)
