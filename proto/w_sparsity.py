import numpy as np, sys
sys.path.insert(0, 'proto'); sys.path.insert(0, '.')
import hqr_multishift as hm
import scipy.linalg as sla
# monkeypatch multishift_sweep to record W sparsity
orig = hm.multishift_sweep
stats = []
def patched(H, Z, ilo, ihi, shifts, w=64, st=None, SP=2):
    n = H.shape[0]; nb = len(shifts); t = 0
    while True:
        p_last = ilo - 1 - SP * (nb - 1) + t
        if p_last > ihi - 2: break
        p_top = max(ilo - 1, p_last); p0 = ilo - 1 + t
        ws = max(ilo, p_top); we = min(ws + w, ihi + 1)
        T = (ihi - 2) - p_last + 1 if we == ihi + 1 else we - 3 - p0
        ww = we - ws
        Hw = H[ws:we, ws:we].copy(); Ww = np.eye(ww, dtype=complex)
        for step in range(T):
            tt = t + step; rots = []
            for i in range(nb):
                p = ilo - 1 - SP * i + tt
                if p < ilo - 1 or p > ihi - 2: continue
                if p == ilo - 1: x = Hw[ilo - ws, ilo - ws] - shifts[i]; y = Hw[ilo + 1 - ws, ilo - ws]
                else: x = Hw[p + 1 - ws, p - ws]; y = Hw[p + 2 - ws, p - ws]
                c, s = hm.givens(x, y); rots.append((p, c, s))
            for (p, c, s) in rots:
                r = p + 1 - ws; hm.rot_rows(Hw, r, c, s, max(p - ws, 0), ww)
                if p >= ilo: Hw[r + 1, p - ws] = 0
            for (p, c, s) in rots:
                k = p + 1 - ws; hm.rot_cols(Hw, k, c, s, 0, min(p + 3, ihi) - ws + 1); hm.rot_cols(Ww, k, c, s, 0, ww)
        if ww == 64:
            nz = (np.abs(Ww) > 0)
            tiles = nz.reshape(8, 8, 8, 8).any(axis=(1, 3))     # [rowtile, coltile]
            # per column tile, k-range in units of 4 rows
            krange = 0
            for J in range(8):
                rows = np.nonzero(nz[:, 8*J:8*J+8].any(axis=1))[0]
                lo = (rows.min() // 4) * 4; hi = (rows.max() // 4 + 1) * 4
                krange += hi - lo
            stats.append((tiles.mean(), krange / (8 * 64), ws == ilo, we == ihi + 1))
        H[ws:we, ws:we] = Hw
        if we < n: H[ws:we, we:n] = Ww.conj().T @ H[ws:we, we:n]
        if ws > 0: H[0:ws, ws:we] = H[0:ws, ws:we] @ Ww
        Z[:, ws:we] = Z[:, ws:we] @ Ww
        t += T
hm.multishift_sweep = patched
from oracle.kbdm_oracle import brain_sim, hankel_matrices, reduce_gep
n = 300
c = brain_sim(2048, 1e-3, 0)
U0, Up1, Up = hankel_matrices(c, n, 1)
A, _, _, _ = reduce_gep(Up1, Up, n)
H, Q = sla.hessenberg(A, calc_q=True)
hm.hqr_multishift(H, Q.copy(), nb=16)
s = np.array(stats)
print("windows", len(s), "mean nonzero tile frac", s[:,0].mean(), "mean k-range frac", s[:,1].mean())
for name, mask in [("intro", s[:,2]==1), ("end", s[:,3]==1), ("middle", (s[:,2]==0)&(s[:,3]==0))]:
    if mask.any(): print(name, mask.sum(), "tile frac", s[mask,0].mean(), "k-range frac", s[mask,1].mean())
