"""Small end-to-end workload touching every kernel family, for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_run.py
(cluster-mode batch, multi-wave ragged batch, rank-deficient member -> Jacobi loop graph, p/q/l options, pooled features, RMSE both
kernels, silhouettes, HDBSCAN core distances + both Prim kernels, batched FID synthesis)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import _native, ensemble, sig_gen, workloads
from llckbdm_b200 import llckbdm as L
c = workloads.brain_sim(512, 1e-3, 1)
r = ensemble.solve_ensemble(c, [96, 33, 64], [96, 20, 64], 2, 1e-3, 5e-4)                       # cluster mode, p = 2, q > 0, l < m
assert (r.status == 0).all()
rng = np.random.default_rng(0)
ms = [int(x) for x in rng.integers(8, 70, 170)]
r = ensemble.solve_ensemble(c, ms, ms, 1, 0.0, 5e-4, chunk=170)                                   # > 148 members: two CTAs per SM in the panels
assert (r.status == 0).all()
r = ensemble.solve_ensemble([workloads.brain_sim(300, 0.0, 0), c], [100, 90], [100, 90], 1, 0.0, 5e-4)   # rank deficient -> Jacobi WHILE graph
assert (r.status == 0).all() and r.info["chunks"][0][14] == 1
r = ensemble.solve_ensemble(c, [40, 50], [40, 50], 1, 0.0, 5e-4, flags=_native.FLAG_NO_GRAPH, options=_native.Options(svd_mode=_native.SVD_JACOBI))
assert (r.status == 0).all()
s, f, st = ensemble.solve_pooled(c, [40, 64, 30], [40, 64, 30], 1, 0.0, 5e-4)
lab = L._fit_all(f, [1, 2, 3])
sil = ensemble.silhouette_samples_device(f, lab)
src, dst, w = ensemble.hdbscan_msts_device(f, [2, 3], single_cta=True)
print("rmse", ensemble.score_candidates(c, 5e-4, [s[:20], workloads.BRAIN_SIM_PARAMS]))
print("rmse long", ensemble.score_candidates(workloads.brain_sim(13000, 1e-3, 2), 5e-4, [workloads.BRAIN_SIM_PARAMS]))
print("fid", sig_gen.multi_fid_batched_device([workloads.BRAIN_SIM_PARAMS, workloads.BRAIN_SIM_PARAMS[:3]], 300, 5e-4).abs().sum().item())
ensemble.release_workspace()
torch.cuda.synchronize()
print("sanitize_run done")
