import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from llckbdm_b200.ensemble import solve_ensemble
from llckbdm_b200.kbdm import kbdm
from oracle.kbdm_oracle import brain_sim, kbdm_oracle, compare_members
c = brain_sim(2048, 1e-3, 7)
ms = [1, 2, 3, 5, 31, 32, 63, 65]
ls = [1, 2, 3, 5, 31, 32, 63, 65]
res = solve_ensemble(c, ms, ls, 1, 0.0, 5e-4)
print("status", res.status.tolist())
for k, (m, l) in enumerate(zip(ms, ls)):
    _, info, mu, D = kbdm_oracle(c, 5e-4, m=m, l=l, return_mu=True)
    dmu, dD = compare_members(res.mu[k, :l], res.D[k, :l], mu, D)
    dsv = np.max(np.abs(res.sing_vals[k, :m] - info.singular_values) / info.singular_values)
    print(f"m={m} l={l} dmu={dmu:.1e} dD={dD:.1e} dsv={dsv:.1e}")
for (m, l, p, q) in [(64, 1, 1, 0.0), (40, 2, 3, 0.0), (100, 7, 1, 1e-2), (33, 33, 4, 0.0)]:
    ll, info = kbdm(c, 5e-4, m=m, l=l, p=p, q=q)
    _, info_o, mu, D = kbdm_oracle(c, 5e-4, m=m, l=l, p=p, q=q, return_mu=True)
    from oracle.kbdm_oracle import mu_from_line_list
    dmu, dD = compare_members(mu_from_line_list(ll, 5e-4), ll[:, 0] * np.exp(1j * ll[:, 3]), mu, D)
    print(f"m={m} l={l} p={p} q={q} dmu={dmu:.1e} dD={dD:.1e}")
# big: N=4096, m=1500 (ld=1536) single member, properties only
c4 = brain_sim(4096, 1e-3, 1)
r = solve_ensemble(c4, [1500], [1500], 1, 0.0, 5e-4)
from oracle.kbdm_oracle import hankel_matrices
U0, _, U1 = hankel_matrices(c4, 1500, 1)
s = np.linalg.svd(U0, compute_uv=False)
print("m=1500 status", r.status.tolist(), "sv rel", np.max(np.abs(r.sing_vals[0] - s) / s), "recon", np.abs((r.D[0][None, :] * r.mu[0][None, :] ** np.arange(32)[:, None]).sum(1) - c4[:32]).max())
