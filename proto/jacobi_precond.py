import numpy as np, sys, scipy.linalg as sla
sys.path.insert(0, '.'); sys.path.insert(0, 'proto')
from oracle.kbdm_oracle import brain_sim, hankel_matrices
from jacobi_svd4 import run
m = int(sys.argv[1]); b = 32
c = brain_sim(2 * m, 1e-3, 0)
U0, _, _ = hankel_matrices(c, m, 1)
Q, R = np.linalg.qr(U0)
Qp, Rp, P = sla.qr(U0, pivoting=True)
for name, X0 in [("plain", U0), ("R^H (no pivot)", R.conj().T), ("R^H (pivoted)", Rp.conj().T), ("R (pivoted)", Rp)]:
    ns, ti, hist = run(X0, b, 1, conv=1e-6)
    print(f"{name:18s} outer sweeps={ns} hist=" + " ".join(f"{h:.1e}" for h in hist), flush=True)
