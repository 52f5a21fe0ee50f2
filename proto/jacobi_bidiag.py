import numpy as np, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'proto')
from oracle.kbdm_oracle import brain_sim, hankel_matrices
from jacobi_svd4 import run

def house(x):
    alpha = x[0]; xn2 = np.sum(np.abs(x[1:]) ** 2)
    if xn2 == 0 and alpha.imag == 0: return np.zeros_like(x), 0.0, alpha
    beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn2), alpha.real)
    tau = (beta - alpha) / beta
    v = x / (alpha - beta); v[0] = 1
    return v, tau, beta

def bidiag(A):
    A = A.copy(); m = A.shape[0]
    for k in range(m):
        v, tau, beta = house(A[k:, k].copy())
        if tau != 0:
            A[k:, k:] -= np.conj(tau) * np.outer(v, v.conj() @ A[k:, k:])
        if k < m - 2:
            v, tau, beta = house(A[k, k + 1:].conj().copy())
            if tau != 0:
                A[k:, k + 1:] -= tau * np.outer(A[k:, k + 1:] @ v, v.conj())     # A <- A P, P = I - tau v v^H built for conj row
    return np.triu(np.tril(A, 1))

m = int(sys.argv[1]); b = 32
c = brain_sim(2 * m, 1e-3, 0)
U0, _, _ = hankel_matrices(c, m, 1)
B = bidiag(U0)
sref = np.linalg.svd(U0, compute_uv=False)
print("bidiag sv err", np.max(np.abs(np.linalg.svd(B, compute_uv=False) - sref) / sref), "offband", np.abs(B - np.triu(np.tril(B,1))).max())
for name, X0 in [("plain U0", U0), ("B", B), ("B^H", B.conj().T.copy())]:
    ns, ti, hist = run(X0, b, 1, conv=1e-6)
    print(f"{name:10s} outer sweeps={ns} hist=" + " ".join(f"{h:.1e}" for h in hist), flush=True)
