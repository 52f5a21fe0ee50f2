import numpy as np, sys, time, scipy.linalg as sla
sys.path.insert(0, '.'); sys.path.insert(0, 'proto')
from oracle.kbdm_oracle import brain_sim, hankel_matrices
from jacobi_svd import block_jacobi_svd
m = int(sys.argv[1]); sigma = float(sys.argv[2]); mode = sys.argv[3]
c = brain_sim(2048, sigma, 0)
U0, _, _ = hankel_matrices(c, m, 1)
sref = np.linalg.svd(U0, compute_uv=False)
if mode == 'qr':
    Q, R = np.linalg.qr(U0); X0 = R.conj().T
elif mode == 'qrp':
    Q, R, P = sla.qr(U0, pivoting=True); X0 = R.conj().T
elif mode == 'qr2':
    Q, R, P = sla.qr(U0, pivoting=True); Q2, R2 = np.linalg.qr(R.conj().T); X0 = R2.conj().T
elif mode == 'lq':
    Q, R = np.linalg.qr(U0); X0 = R   # jacobi on R itself (upper triangular columns)
else:
    X0 = U0
X, s, V, ns = block_jacobi_svd(X0, b=32, tol=1e-13, max_sweeps=14)
ss = np.sort(s)[::-1][:m]
print(mode, "sweeps", ns, "max rel sv err", np.max(np.abs(ss - sref) / sref))
