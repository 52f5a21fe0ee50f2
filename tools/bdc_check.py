"""Stage check of the divide-and-conquer SVD of real bidiagonals: python tools/bdc_check.py [m1 m2 ...]
Random bidiagonals (graded, clustered, with zeros) of ragged sizes in ONE batch, compared with numpy.linalg.svd."""
import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import _native


def run_bdc(ds, es):
    lib = _native.load()
    batch = len(ds)
    ms = [len(d) for d in ds]
    ld = lib.llck_leading_dim(max(ms))
    dev = torch.device("cuda:0")
    D = np.zeros((batch, ld)); E = np.zeros((batch, ld))
    for b in range(batch):
        D[b, :ms[b]] = ds[b]; E[b, :ms[b] - 1] = es[b]
    Dd = torch.from_numpy(D).to(dev); Ed = torch.from_numpy(E).to(dev)
    sv = torch.zeros((batch, ld), dtype=torch.float64, device=dev)
    Us = torch.zeros((batch, ld, ld, 2), dtype=torch.float64, device=dev); V = torch.zeros_like(Us)
    fb = torch.zeros(batch, dtype=torch.int32, device=dev)
    marr = (ctypes.c_int32 * batch)(*ms)
    rc = lib.llck_bdc_test(Dd.data_ptr(), Ed.data_ptr(), marr, batch, ld, sv.data_ptr(), Us.data_ptr(), V.data_ptr(), fb.data_ptr(), None)
    assert rc == 0, rc
    svh = sv.cpu().numpy(); Ush = Us.cpu().numpy()[..., 0]; Vh = V.cpu().numpy()[..., 0]; fbh = fb.cpu().numpy()
    out = []
    for b in range(batch):
        m = ms[b]
        out.append((svh[b, :m], Ush[b].T[:m, :m], Vh[b].T[:m, :m], int(fbh[b])))   # column-major -> [row, col]
    return out


def check(ds, es, names):
    res = run_bdc(ds, es)
    worst = 0.0
    for (s, Us, V, fb), d, e, name in zip(res, ds, es, names):
        m = len(d)
        B = np.diag(d) + np.diag(e, 1)
        sref = np.linalg.svd(B, compute_uv=False)
        if fb:
            print(f"{name:28s} m={m:5d} FALLBACK flagged (smin/smax={sref[-1] / sref[0]:.1e})")
            continue
        U = Us / s
        es_ = np.max(np.abs(s - sref)) / sref[0]
        ou = np.abs(U.T @ U - np.eye(m)).max(); ov = np.abs(V.T @ V - np.eye(m)).max()
        rs = np.abs(B - Us @ V.T).max() / sref[0]
        worst = max(worst, es_, ou, ov, rs)
        print(f"{name:28s} m={m:5d} sv abs/smax={es_:.1e} orthU={ou:.1e} orthV={ov:.1e} resid={rs:.1e} smin/smax={sref[-1] / sref[0]:.1e}")
    return worst


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 17, 32, 33, 64, 65, 96, 130, 257, 300, 512, 700, 1000, 1024]
    rng = np.random.default_rng(0)
    ds, es, names = [], [], []
    for m in sizes:
        ds.append(rng.standard_normal(m)); es.append(rng.standard_normal(m - 1)); names.append("gaussian")
    for m in sizes[-6:]:
        g = np.logspace(0, -5, m)
        ds.append(g * rng.standard_normal(m)); es.append(g[:-1] * rng.standard_normal(m - 1)); names.append("graded 1e-5")
        ds.append(np.ones(m)); es.append(np.full(m - 1, 1e-3)); names.append("clustered (1, 1e-3)")
        d = rng.standard_normal(m); e = rng.standard_normal(m - 1); e[m // 3] = 0.0; e[m // 2] = 1e-18
        ds.append(d); es.append(e); names.append("split (zero e)")
        ds.append(np.arange(1, m + 1, dtype=float)); es.append(np.zeros(m - 1)); names.append("diagonal")
    w = check(ds, es, names)
    print("worst:", w)
    assert w < 1e-11, w
    print("BDC_OK")
