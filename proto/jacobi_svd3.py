import numpy as np, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'proto')
from oracle.kbdm_oracle import brain_sim, hankel_matrices
from jacobi_svd import rr_schedule
from herm_jacobi import herm_jacobi

def block_jacobi_svd(A, b=32, tol=1e-14, max_sweeps=30, verbose=True):
    m = A.shape[0]
    nb = (m + b - 1) // b
    if nb % 2: nb += 1
    mp = nb * b
    X = np.zeros((m, mp), dtype=complex); X[:, :m] = A
    V = np.eye(mp, dtype=complex)
    rounds = rr_schedule(nb)
    tot_inner = 0
    for sweep in range(max_sweeps):
        maxoff = 0.0; nrot = 0; inner = 0
        for pairs in rounds:
            for (i, j) in pairs:
                cols = np.r_[i*b:(i+1)*b, j*b:(j+1)*b]
                Xp = X[:, cols]
                G = Xp.conj().T @ Xp
                d = np.sqrt(np.abs(np.diag(G)).clip(1e-300))
                off = np.abs(G - np.diag(np.diag(G))) / (d[:, None] * d[None, :])
                mo = off.max()
                maxoff = max(maxoff, mo)
                if mo < tol: continue
                nrot += 1
                w, J, nsw = herm_jacobi(G, tol=tol/4)
                inner += nsw
                o = np.argsort(-w); J = J[:, o]
                X[:, cols] = Xp @ J
                V[:, cols] = V[:, cols] @ J
        tot_inner += inner
        if verbose: print(f"sweep {sweep}: maxoff={maxoff:.3e} blockrots={nrot} avg inner sweeps={inner/max(nrot,1):.1f}")
        if maxoff < tol: break
    s = np.linalg.norm(X, axis=0)
    return X, s, V, sweep + 1

if __name__ == '__main__':
    m = int(sys.argv[1]); sigma = float(sys.argv[2]); b = int(sys.argv[3])
    c = brain_sim(2048, sigma, 0)
    U0, _, _ = hankel_matrices(c, m, 1)
    X, s, V, ns = block_jacobi_svd(U0, b=b)
    sref = np.linalg.svd(U0, compute_uv=False)
    o = np.argsort(-s)[:m]
    print("sweeps", ns, "max rel sv err", np.max(np.abs(s[o] - sref) / sref))
    L = X[:, o] / s[o]; R = V[:m, o]
    print("orth L", np.abs(L.conj().T @ L - np.eye(m)).max(), "orth R", np.abs(R.conj().T@R - np.eye(m)).max(), "recon", np.abs((L*s[o]) @ R.conj().T - U0).max())
