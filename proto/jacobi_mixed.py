"""Mixed-precision block Jacobi: k sweeps in complex64, then V re-orthonormalised in FP64 (Newton-Schulz),
X = U0 V recomputed in FP64, then FP64 sweeps to convergence.  Counts the FP64 sweeps needed."""
import numpy as np, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'proto')
from oracle.kbdm_oracle import brain_sim, hankel_matrices
from jacobi_svd import rr_schedule
from herm_jacobi import herm_jacobi

def sweeps(X, V, b, nsweeps, dtype, conv, inner_max=1, tol=1e-14):
    m, mp = X.shape
    nb = mp // b
    rounds = rr_schedule(nb)
    hist = []
    for sweep in range(nsweeps):
        maxoff = 0.0
        for pairs in rounds:
            for (i, j) in pairs:
                cols = np.r_[i*b:(i+1)*b, j*b:(j+1)*b]
                Xp = X[:, cols]
                G = (Xp.conj().T @ Xp).astype(np.complex128)
                d = np.sqrt(np.abs(np.diag(G)).clip(1e-300))
                mo = (np.abs(G - np.diag(np.diag(G))) / (d[:, None] * d[None, :])).max()
                maxoff = max(maxoff, mo)
                if mo < tol: continue
                cap = inner_max if mo > 1e-3 else max(inner_max, 2)
                w, J, nsw = herm_jacobi(G, max_sweeps=cap, tol=tol/4)
                o = np.argsort(-w); J = J[:, o].astype(dtype)
                X[:, cols] = Xp @ J
                V[:, cols] = V[:, cols] @ J
        hist.append(maxoff)
        if maxoff < conv: break
    return hist

m = int(sys.argv[1]); k32 = int(sys.argv[2]); b = 32
c = brain_sim(2 * m, 1e-3, 0)
U0, _, _ = hankel_matrices(c, m, 1)
X = U0.astype(np.complex64); V = np.eye(m, dtype=np.complex64)
h32 = sweeps(X, V, b, k32, np.complex64, conv=1e-5)
print("fp32 hist:", " ".join(f"{h:.1e}" for h in h32), flush=True)
V64 = V.astype(np.complex128)
for it in range(3):
    E = V64.conj().T @ V64
    print("  orth err", np.abs(E - np.eye(m)).max())
    V64 = V64 @ (1.5 * np.eye(m) - 0.5 * E)
X64 = U0 @ V64
h64 = sweeps(X64, V64, b, 20, np.complex128, conv=1e-6)
print("fp64 hist:", " ".join(f"{h:.1e}" for h in h64), " -> fp64 sweeps:", len(h64), flush=True)
s = np.sort(np.linalg.norm(X64, axis=0))[::-1]
sref = np.linalg.svd(U0, compute_uv=False)
print("sv rel err", np.max(np.abs(s - sref) / sref), "V orth", np.abs(V64.conj().T @ V64 - np.eye(m)).max(), "XV resid", np.abs(U0 @ V64 - X64).max())
