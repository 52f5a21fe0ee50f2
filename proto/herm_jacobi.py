import numpy as np
from jacobi_svd import rr_schedule

def herm_jacobi(G, max_sweeps=12, tol=1e-15):
    """two-sided cyclic Jacobi (parallel round-robin ordering) for Hermitian G; returns (w, J) with G = J diag(w) J^H"""
    n = G.shape[0]
    G = G.copy(); J = np.eye(n, dtype=complex)
    rounds = rr_schedule(n)
    nsw = 0
    for sweep in range(max_sweeps):
        rotated = False
        for pairs in rounds:
            p = np.array([a for a, b in pairs]); q = np.array([b for a, b in pairs])
            gpp = G[p, p].real; gqq = G[q, q].real; gpq = G[p, q]
            ab = np.abs(gpq)
            act = ab > tol * np.sqrt(np.abs(gpp * gqq))
            if not act.any(): continue
            rotated = True
            # rotation: [c, s; -conj(s), c] columns
            with np.errstate(divide='ignore', invalid='ignore'):
                zeta = (gqq - gpp) / (2 * ab)
                t = np.sign(zeta) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta))
                t = np.where(zeta == 0, 1.0, t)
                c = 1 / np.sqrt(1 + t * t)
                s = c * t * gpq / ab           # complex s with phase of gpq
            c = np.where(act, c, 1.0); s = np.where(act, s, 0.0)
            # columns: new_p = c*colp - conj(s)*colq ; new_q = s*colp + c*colq
            Gp = G[:, p].copy(); Gq = G[:, q].copy()
            G[:, p] = c * Gp - np.conj(s) * Gq
            G[:, q] = s * Gp + c * Gq
            Gp = G[p, :].copy(); Gq = G[q, :].copy()
            G[p, :] = c[:, None] * Gp - s[:, None] * Gq
            G[q, :] = np.conj(s)[:, None] * Gp + c[:, None] * Gq
            Jp = J[:, p].copy(); Jq = J[:, q].copy()
            J[:, p] = c * Jp - np.conj(s) * Jq
            J[:, q] = s * Jp + c * Jq
        nsw += 1
        if not rotated: break
    return G.diagonal().real.copy(), J, nsw

if __name__ == '__main__':
    rng = np.random.default_rng(0)
    X = rng.standard_normal((200, 64)) + 1j * rng.standard_normal((200, 64))
    G = X.conj().T @ X
    w, J, nsw = herm_jacobi(G)
    print(nsw, np.abs(J.conj().T @ G @ J - np.diag(w)).max(), np.abs(J.conj().T @ J - np.eye(64)).max())
