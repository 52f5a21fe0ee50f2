"""BASELINE.json configs C3 / C4 / C5 (scaled subsets) through the public API: timing + sampled parity vs the oracle.

    python tools/configs_check.py [n_voxels_c4] [n_members_c5]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch  # noqa: E402

from llckbdm_b200.ensemble import solve_ensemble  # noqa: E402
from llckbdm_b200.min_rmse_kbdm import min_rmse_kbdm  # noqa: E402
from oracle.kbdm_oracle import BRAIN_SIM_PARAMS, brain_sim, compare_members, kbdm_oracle, multi_fid_oracle  # noqa: E402

DWELL = 5e-4
out = {}

# ---- C3: min_rmse_kbdm sweep over m on a 4096-pt FID with pseudo-noise ----
c = brain_sim(4096, 1e-3, 0)
rng = np.random.default_rng(1)
c3 = c + 1e-6 * (rng.standard_normal(4096) + 1j * rng.standard_normal(4096))
m_range = list(range(256, 1025, 64))
torch.cuda.synchronize(); t0 = time.perf_counter()
r = min_rmse_kbdm(c3, DWELL, m_range=m_range, l=None)
torch.cuda.synchronize(); t1 = time.perf_counter()
out["c3_min_rmse"] = {"members": len(m_range), "seconds": t1 - t0, "min_index": int(r.min_index), "min_rmse": float(r.min_rmse),
                      "n_samples": len(r.samples)}
print("C3", out["c3_min_rmse"], flush=True)

# ---- C4: MRSI voxels, N=1024, m=l=512, per-voxel perturbed parameters + noise ----
nvox = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t = np.linspace(0, DWELL * 1024, 1024, endpoint=False)
sigs = []
for v in range(nvox):
    g = np.random.default_rng(v)
    p = BRAIN_SIM_PARAMS.copy()
    p[:, 0] *= g.uniform(0.5, 1.5, 16); p[:, 1] *= g.uniform(0.8, 1.2, 16); p[:, 2] += g.normal(0, 2.0, 16)
    sigs.append(multi_fid_oracle(t, p) + 1e-3 * (g.standard_normal(1024) + 1j * g.standard_normal(1024)))
torch.cuda.synchronize(); t0 = time.perf_counter()
res = solve_ensemble(sigs, [512] * nvox, [512] * nvox, 1, 0.0, DWELL)
torch.cuda.synchronize(); t1 = time.perf_counter()
worst = 0.0
for v in (0, nvox // 2, nvox - 1):
    _, _, mu, D = kbdm_oracle(sigs[v], DWELL, m=512, return_mu=True)
    dmu, dD = compare_members(res.mu[v], res.D[v], mu, D)
    worst = max(worst, dmu, dD)
out["c4_mrsi"] = {"voxels": nvox, "seconds": t1 - t0, "voxels_per_s": nvox / (t1 - t0), "bad_status": int((res.status != 0).sum()),
                  "worst_rel_err_vs_oracle_3_voxels": worst, "extrapolated_65536_voxels_s": 65536 * (t1 - t0) / nvox}
print("C4", out["c4_mrsi"], flush=True)

# ---- C5: large LLC ensemble subset: m_k = 512 + (k mod 513), pseudo-noise seed 1000+k on a 4096-pt FID ----
nmem = int(sys.argv[2]) if len(sys.argv) > 2 else 296
sig5, ms = [], []
for k in range(nmem):
    g = np.random.default_rng(1000 + k)
    sig5.append(c + 1e-6 * (g.standard_normal(4096) + 1j * g.standard_normal(4096)))
    ms.append(512 + (k * 37) % 513)          # spread over [512, 1024] (k mod 513 would only reach 512+nmem)
torch.cuda.synchronize(); t0 = time.perf_counter()
res5 = solve_ensemble(sig5, ms, ms, 1, 0.0, DWELL)
torch.cuda.synchronize(); t1 = time.perf_counter()
k = int(np.argmin(ms))
_, _, mu, D = kbdm_oracle(sig5[k], DWELL, m=ms[k], return_mu=True)
dmu, dD = compare_members(res5.mu[k, :ms[k]], res5.D[k, :ms[k]], mu, D)
out["c5_large_ensemble"] = {"members": nmem, "m_min": min(ms), "m_max": max(ms), "seconds": t1 - t0, "members_per_s": nmem / (t1 - t0),
                            "bad_status": int((res5.status != 0).sum()), "rel_err_vs_oracle_smallest_member": max(dmu, dD)}
print("C5", out["c5_large_ensemble"], flush=True)
print(json.dumps(out))
