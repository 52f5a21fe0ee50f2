import numpy as np, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'proto')
from oracle.kbdm_oracle import brain_sim, hankel_matrices
from jacobi_svd import rr_schedule
from herm_jacobi import herm_jacobi

def run(A, b, inner_max, tol=1e-14, conv=1e-7, max_sweeps=20, sort=True):
    m = A.shape[0]
    nb = (m + b - 1) // b
    if nb % 2: nb += 1
    mp = nb * b
    X = np.zeros((m, mp), dtype=complex); X[:, :m] = A
    rounds = rr_schedule(nb)
    tot_inner = 0; hist = []
    for sweep in range(max_sweeps):
        maxoff = 0.0; inner = 0
        for pairs in rounds:
            for (i, j) in pairs:
                cols = np.r_[i*b:(i+1)*b, j*b:(j+1)*b]
                Xp = X[:, cols]
                G = Xp.conj().T @ Xp
                d = np.sqrt(np.abs(np.diag(G)).clip(1e-300))
                mo = (np.abs(G - np.diag(np.diag(G))) / (d[:, None] * d[None, :])).max()
                maxoff = max(maxoff, mo)
                if mo < tol: continue
                w, J, nsw = herm_jacobi(G, max_sweeps=inner_max, tol=tol/4)
                inner += nsw
                if sort:
                    o = np.argsort(-w); J = J[:, o]
                X[:, cols] = Xp @ J
        tot_inner += inner; hist.append(maxoff)
        if maxoff < conv: break
    return sweep + 1, tot_inner, hist

if __name__ == '__main__':
    m = int(sys.argv[1]); b = int(sys.argv[2])
    c = brain_sim(2048, 1e-3, 0)
    U0, _, _ = hankel_matrices(c, m, 1)
    for inner_max in [1, 2, 3, 12]:
        ns, ti, hist = run(U0, b, inner_max)
        print(f"m={m} b={b} inner_max={inner_max}: outer sweeps={ns} total inner sweeps={ti} hist=" + " ".join(f"{h:.1e}" for h in hist))
