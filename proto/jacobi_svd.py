import numpy as np, sys, time
sys.path.insert(0, '.')
from oracle.kbdm_oracle import brain_sim, hankel_matrices

def rr_schedule(nb):
    """round-robin tournament: nb even -> nb-1 rounds of nb/2 disjoint pairs"""
    idx = list(range(nb))
    rounds = []
    for r in range(nb - 1):
        pairs = [(min(idx[i], idx[nb-1-i]), max(idx[i], idx[nb-1-i])) for i in range(nb // 2)]
        rounds.append(pairs)
        idx = [idx[0]] + [idx[-1]] + idx[1:-1]
    return rounds

def block_jacobi_svd(A, b=32, tol=1e-15, max_sweeps=30, sort_inner=True, verbose=True):
    m = A.shape[0]
    nb = (m + b - 1) // b
    if nb % 2: nb += 1
    mp = nb * b
    X = np.zeros((m, mp), dtype=complex); X[:, :m] = A
    V = np.eye(mp, dtype=complex)
    rounds = rr_schedule(nb)
    for sweep in range(max_sweeps):
        maxoff = 0.0; nrot = 0
        for pairs in rounds:
            for (i, j) in pairs:
                cols = np.r_[i*b:(i+1)*b, j*b:(j+1)*b]
                Xp = X[:, cols]
                G = Xp.conj().T @ Xp
                d = np.sqrt(np.abs(np.diag(G)).clip(1e-300))
                off = np.abs(G - np.diag(np.diag(G))) / (d[:, None] * d[None, :])
                # ignore zero columns
                off[np.isnan(off)] = 0
                mo = off.max()
                maxoff = max(maxoff, mo)
                if mo < tol: continue
                nrot += 1
                w, J = np.linalg.eigh(G)
                J = J[:, ::-1]  # descending
                X[:, cols] = Xp @ J
                V[:, cols] = V[:, cols] @ J
        if verbose: print(f"sweep {sweep}: maxoff={maxoff:.3e} blockrots={nrot}")
        if maxoff < tol: break
    s = np.linalg.norm(X, axis=0)
    return X, s, V, sweep + 1

if __name__ == '__main__':
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
    b = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    c = brain_sim(2048, sigma, 0)
    U0, _, _ = hankel_matrices(c, m, 1)
    t0 = time.time()
    X, s, V, ns = block_jacobi_svd(U0, b=b, tol=1e-14)
    print("time", time.time() - t0)
    sref = np.linalg.svd(U0, compute_uv=False)
    ss = np.sort(s)[::-1][:m]
    print("sweeps", ns, "max rel sv err", np.max(np.abs(ss - sref) / sref), "cond", sref[0]/sref[-1])
