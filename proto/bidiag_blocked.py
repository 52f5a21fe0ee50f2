import numpy as np, sys
def house(x):
    alpha = x[0]; xn2 = np.sum(np.abs(x[1:]) ** 2)
    if xn2 == 0 and alpha.imag == 0:
        v = np.zeros_like(x); v[0] = 1
        return v, 0.0, alpha.real
    beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn2), alpha.real)
    tau = (beta - alpha) / beta
    v = x / (alpha - beta); v[0] = 1
    return v, tau, beta

def bidiag_blocked(A0, nb=32):
    A = A0.copy(); m = A.shape[0]
    d = np.zeros(m); e = np.zeros(m - 1)
    panels = []
    for k0 in range(0, m, nb):
        nbp = min(nb, m - k0)
        V = np.zeros((m, nbp), dtype=complex); Y = np.zeros((m, nbp), dtype=complex)
        X = np.zeros((m, nbp), dtype=complex); U = np.zeros((m, nbp), dtype=complex)
        TQ = np.zeros((nbp, nbp), dtype=complex); TP = np.zeros((nbp, nbp), dtype=complex)
        for i in range(nbp):
            c = k0 + i
            a = A[:, c] - V[:, :i] @ np.conj(Y[c, :i]) - X[:, :i] @ np.conj(U[c, :i])
            v = np.zeros(m, dtype=complex)
            vv, tau, beta = house(a[c:].copy())
            v[c:] = vv; d[c] = beta
            zv = V[:, :i].conj().T @ v; zx = X[:, :i].conj().T @ v
            TQ[:i, i] = -tau * (TQ[:i, :i] @ zv); TQ[i, i] = tau
            V[:, i] = v
            if c < m - 1:
                y = np.zeros(m, dtype=complex)
                y[c + 1:] = tau * (A[c:, c + 1:].conj().T @ v[c:] - Y[c + 1:, :i] @ zv - U[c + 1:, :i] @ zx)
                Y[:, i] = y
                r = A[c, :] - V[c, :i + 1] @ Y[:, :i + 1].conj().T - X[c, :i] @ U[:, :i].conj().T      # row c of (A_i - v y^H)
                u = np.zeros(m, dtype=complex)
                uu, pi_, betap = house(np.conj(r[c + 1:]).copy())
                u[c + 1:] = uu; e[c] = betap
                zy = Y[:, :i + 1].conj().T @ u; zu = U[:, :i].conj().T @ u
                TP[:i, i] = -pi_ * (TP[:i, :i] @ zu); TP[i, i] = pi_
                U[:, i] = u
                x = np.zeros(m, dtype=complex)
                x[c + 1:] = pi_ * (A[c + 1:, c + 1:] @ u[c + 1:] - V[c + 1:, :i + 1] @ zy - X[c + 1:, :i] @ zu)
                X[:, i] = x
        e_ = k0 + nbp
        if e_ < m:
            A[e_:, e_:] -= V[e_:, :] @ Y[e_:, :].conj().T + X[e_:, :] @ U[e_:, :].conj().T
        panels.append((k0, nbp, V, TQ, U, TP))
    # Q, P by blocked backward accumulation
    Q = np.eye(m, dtype=complex); P = np.eye(m, dtype=complex)
    for (k0, nbp, V, TQ, U, TP) in reversed(panels):
        Q[k0:, k0:] -= V[k0:, :] @ (TQ @ (V[k0:, :].conj().T @ Q[k0:, k0:]))
        P[k0:, k0:] -= U[k0:, :] @ (TP @ (U[k0:, :].conj().T @ P[k0:, k0:]))
    return d, e, Q, P

if __name__ == '__main__':
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = np.random.default_rng(0)
    A = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
    d, e, Q, P = bidiag_blocked(A)
    B = np.diag(d) + np.diag(e, 1)
    print("resid", np.abs(Q @ B @ P.conj().T - A).max(), "Qorth", np.abs(Q.conj().T @ Q - np.eye(m)).max(), "Porth", np.abs(P.conj().T @ P - np.eye(m)).max())
