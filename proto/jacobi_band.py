import numpy as np, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'proto')
from oracle.kbdm_oracle import brain_sim, hankel_matrices
from jacobi_svd4 import run
def band_reduce(A, nb=32):
    A = A.copy(); m = A.shape[0]
    for k in range(0, m, nb):
        e = min(k + nb, m)
        Q, R = np.linalg.qr(A[k:, k:e], mode='complete')        # left: zero below diagonal of block column
        A[k:, k:] = Q.conj().T @ A[k:, k:]
        if e < m - 1:
            # right: LQ of block row A[k:e, e:] -> zero right of the band
            Q2, R2 = np.linalg.qr(A[k:e, e:].conj().T, mode='complete')
            A[k:, e:] = A[k:, e:] @ Q2
    return A
m = int(sys.argv[1]); b = 32
c = brain_sim(2 * m, 1e-3, 0)
U0, _, _ = hankel_matrices(c, m, 1)
Bd = band_reduce(U0, 32)
mask = np.triu(np.ones((m, m)), 0) * np.tril(np.ones((m, m)), 2 * 32)
print("outside band", np.abs(Bd * (1 - mask)).max(), "sv err", np.max(np.abs(np.linalg.svd(Bd, compute_uv=False) - np.linalg.svd(U0, compute_uv=False))))
Q, R = np.linalg.qr(U0)
for name, X0 in [("band^H", Bd.conj().T.copy()), ("band", Bd), ("R^H", R.conj().T.copy())]:
    ns, ti, hist = run(X0, b, 1, conv=1e-6)
    print(f"{name:10s} outer sweeps={ns} hist=" + " ".join(f"{h:.1e}" for h in hist), flush=True)
