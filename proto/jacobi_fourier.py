import numpy as np, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'proto')
from oracle.kbdm_oracle import brain_sim, hankel_matrices
from jacobi_svd4 import run
m = int(sys.argv[1]); b = 32
c = brain_sim(2048, 1e-3, 0)
U0, _, _ = hankel_matrices(c, m, 1)
F = np.fft.fft(np.eye(m)) / np.sqrt(m)
for name, X0 in [("plain", U0), ("fourier", U0 @ F), ("fourier-conj", U0 @ F.conj())]:
    G = X0.conj().T @ X0
    d = np.sqrt(np.diag(G).real)
    off = np.abs(G - np.diag(np.diag(G))) / np.outer(d, d)
    print(name, "initial scaled off: max", off.max(), "mean", off.mean(), "frac>0.1:", (off > 0.1).mean())
    ns, ti, hist = run(X0, b, 2)
    print(f"   outer sweeps={ns} total inner sweeps={ti} hist=" + " ".join(f"{h:.1e}" for h in hist))
