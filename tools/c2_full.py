"""Config C2 end to end: LLC-KBDM on the brain_sim FID, 100 truncations m in [700, 1024] + clustering + selection.
Prints the latency of each stage of llc_kbdm (reference llckbdm.py:41-141).   python tools/c2_full.py [members]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import llckbdm as L
from llckbdm_b200.min_rmse_kbdm import min_rmse_kbdm
from llckbdm_b200.sampling import filter_samples, sample_kbdm
from oracle.kbdm_oracle import brain_sim

M = int(sys.argv[1]) if len(sys.argv) > 1 else 100
c = brain_sim(2048, 1e-3, 0)
m_range = [700 + round(k * 324 / max(M - 1, 1)) for k in range(M)]
sample_kbdm(c, 5e-4, m_range[:2], p=1, l=None)          # warm-up (library load, allocator)
torch.cuda.synchronize()
t = {}
t0 = time.perf_counter(); lls, _ = sample_kbdm(c, 5e-4, m_range, p=1, l=None); torch.cuda.synchronize(); t["solve_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); samples = filter_samples(np.concatenate(lls)); feats = L._transform_line_lists(samples, 5e-4); t["pool_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); labelings = L._fit_all(feats, list(range(1, M))); t["hdbscan_fits_s"] = time.perf_counter() - t0
if L._gpu_fit_supported(feats, list(range(1, M))):      # split of the line above: device spanning trees vs host tree condensation
    from llckbdm_b200.ensemble import hdbscan_msts_device
    t0 = time.perf_counter(); hdbscan_msts_device(feats, list(range(1, M))); torch.cuda.synchronize(); t["of_which_device_core_and_mst_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); results = L._results_from_labelings(samples, feats, labelings); t["silhouette_summarise_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); best = min_rmse_kbdm(c, 5e-4, samples=[r.summarized_line_list for r in results]); t["rmse_select_s"] = time.perf_counter() - t0
t["total_s"] = sum(v for k, v in t.items() if not k.startswith("of_which"))
t.update(members=M, pooled_points=int(len(samples)), clusterings=len(results), host_cores=os.cpu_count(),
         best_clusters=int(len(best.line_list)), best_rmse=float(best.min_rmse))
print("C2_FULL " + json.dumps(t))
