"""Time the clustering-stage kernels at the C2 size: python tools/mst_time.py [n] [fits]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import _native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 43212
F = int(sys.argv[2]) if len(sys.argv) > 2 else 99
lib = _native.load()
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
cent = rng.uniform(-1, 1, (400, 3))
X = np.concatenate([np.repeat(cent, 60, axis=0) + 1e-3 * rng.standard_normal((24000, 3)), rng.uniform(-1, 1, (n - 24000, 3))])
X = np.column_stack([X, np.zeros(len(X))])
Xd = torch.from_numpy(X).to(dev)
kmax = F + 1
core = torch.empty((kmax, n), dtype=torch.float64, device=dev)
rows = torch.arange(1, F + 1, dtype=torch.int32, device=dev)
mr = torch.empty((F, n), dtype=torch.float64, device=dev); cs = torch.empty((F, n), dtype=torch.int32, device=dev)
src = torch.empty((F, n - 1), dtype=torch.int64, device=dev); dst = torch.empty_like(src); w = torch.empty((F, n - 1), dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
def timed(fn):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
for rep in range(2):
    t_core = timed(lambda: lib.llck_hdbscan_core_distances(Xd.data_ptr(), n, kmax, core.data_ptr(), st))
    t_cl = timed(lambda: lib.llck_hdbscan_mst(Xd.data_ptr(), n, core.data_ptr(), rows.data_ptr(), F, mr.data_ptr(), cs.data_ptr(), src.data_ptr(), dst.data_ptr(), w.data_ptr(), 2, st))      # 2 = LLCK_MST_DIM3: the 4th coordinate is 0 for every point
    a = (src.clone(), dst.clone(), w.clone())
    t_1 = timed(lambda: lib.llck_hdbscan_mst(Xd.data_ptr(), n, core.data_ptr(), rows.data_ptr(), F, mr.data_ptr(), cs.data_ptr(), src.data_ptr(), dst.data_ptr(), w.data_ptr(), 1, st))
    same = bool((a[0] == src).all() and (a[1] == dst).all() and (a[2] == w).all())
    t0 = time.perf_counter(); h = (src.cpu(), dst.cpu(), w.cpu()); t_d2h = time.perf_counter() - t0
    print(f"n={n} fits={F}: core distances {t_core:.1f} ms | Prim cluster kernel {t_cl:.1f} ms | Prim single-CTA kernel {t_1:.1f} ms | identical={same} | D2H of the edges {t_d2h*1e3:.1f} ms", flush=True)
