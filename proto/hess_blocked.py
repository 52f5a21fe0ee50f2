import numpy as np, sys
def hess_blocked(A, nb=32):
    A = A.copy(); n = A.shape[0]
    taus = np.zeros(n, dtype=complex); Ts = []
    k0 = 0
    while k0 < n - 2:
        nbp = min(nb, n - 2 - k0)
        V = np.zeros((n, nbp), dtype=complex); T = np.zeros((nbp, nbp), dtype=complex); Y = np.zeros((n, nbp), dtype=complex)
        for j in range(nbp):
            c = k0 + j
            b = A[:, c] - Y[:, :j] @ np.conj(V[c, :j])
            b = b - V[:, :j] @ (T[:j, :j].conj().T @ (V[:, :j].conj().T @ b))
            alpha = b[c + 1]; xn2 = np.sum(np.abs(b[c + 2:]) ** 2)
            v = np.zeros(n, dtype=complex)
            if xn2 == 0 and alpha.imag == 0:
                tau = 0.0; v[c + 1] = 1.0
            else:
                beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn2), alpha.real)
                tau = (beta - alpha) / beta
                v[c + 1] = 1.0; v[c + 2:] = b[c + 2:] / (alpha - beta)
                b[c + 1] = beta; b[c + 2:] = 0
            A_col_store = b.copy(); A_col_store[c + 2:] = v[c + 2:]
            z = V[:, :j].conj().T @ v
            T[:j, j] = -tau * (T[:j, :j] @ z); T[j, j] = tau
            y = A[:, c + 1:] @ v[c + 1:]          # big gemv with panel-start A (cols > c untouched so far)
            Y[:, j] = tau * (y - Y[:, :j] @ z)
            V[:, j] = v
            A[:, c] = A_col_store
            taus[c] = tau
        e = k0 + nbp
        A[:, e:] -= Y @ V[e:, :].conj().T
        C = A[k0 + 1:, e:]
        Vs = V[k0 + 1:, :]
        A[k0 + 1:, e:] = C - Vs @ (T.conj().T @ (Vs.conj().T @ C))
        Ts.append((k0, nbp, T))
        k0 = e
    # Q formation (blocked, backward)
    Q = np.eye(n, dtype=complex)
    for (k0, nbp, T) in reversed(Ts):
        V = np.zeros((n, nbp), dtype=complex)
        for j in range(nbp):
            c = k0 + j; V[c + 1, j] = 1.0; V[c + 2:, j] = A[c + 2:, c]
        Vs = V[k0 + 1:, :]
        Qs = Q[k0 + 1:, k0 + 1:]
        Q[k0 + 1:, k0 + 1:] = Qs - Vs @ (T @ (Vs.conj().T @ Qs))
    H = np.triu(A, -1)
    return H, Q
if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = np.random.default_rng(0)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H, Q = hess_blocked(A, 32)
    print("resid", np.abs(Q @ H @ Q.conj().T - A).max(), "orth", np.abs(Q.conj().T @ Q - np.eye(n)).max())
