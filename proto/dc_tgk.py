"""Prototype: SVD of a real upper-bidiagonal B (d, e) by divide & conquer on the Golub-Kahan tridiagonal
T = shuffle([[0, B^T], [B, 0]]) (zero diagonal, off-diagonals d1,e1,d2,e2,...,dn).

Cuppen's rank-one tearing + secular equation with origin shifts + Gu/Eisenstat z-hat (Loewner) for orthogonality.
Written the way the CUDA kernels are organised (level by level, explicit deflation/permutation arrays).

usage: python proto/dc_tgk.py m [sigma]
"""
import sys
import numpy as np

EPS = np.finfo(float).eps
USE_BNS = True


def leaf_eig(a, b):
    T = np.diag(a) + np.diag(b, 1) + np.diag(b, -1)
    w, Q = np.linalg.eigh(T)
    return w, Q


def secular_roots(dd, z, rho):
    """Roots of 1 + rho * sum z_i^2 / (dd_i - lam) for ascending dd, rho > 0, all z_i != 0.
    Returns (org, mu): lam_j = dd[org_j] + mu_j with the origin the nearer pole."""
    k = len(dd)
    z2 = z * z
    # interval j: (dd_j, dd_{j+1}), last: (dd_k, dd_k + rho*sum z2)
    upper = np.append(dd[1:], dd[-1] + rho * z2.sum())
    gap = upper - dd
    mid = dd + 0.5 * gap
    # f(mid) vectorised with differences taken against dd_j (exact-ish)
    delta_mid = (dd[None, :] - dd[:, None]) - 0.5 * gap[:, None]      # [j, i] = dd_i - mid_j
    fmid = 1.0 + rho * (z2[None, :] / delta_mid).sum(axis=1)
    left = fmid > 0          # root in the left half: origin dd_j, mu in (0, gap/2]
    left[-1] = True          # last root: always origin dd_k
    org = np.where(left, np.arange(k), np.minimum(np.arange(k) + 1, k - 1))
    delta = dd[None, :] - dd[org][:, None]                              # [j, i] = dd_i - origin_j
    lo = np.where(left, 0.0, -0.5 * gap)
    hi = np.where(left, 0.5 * gap, 0.0)
    hi[-1] = gap[-1]
    lo = lo.copy(); hi = hi.copy()

    def g(mu):
        return 1.0 + rho * (z2[None, :] / (delta - mu[:, None])).sum(axis=1)

    # bisection on mu (g increasing in mu); geometric midpoint when the bracket spans orders of magnitude
    for it in range(200):
        same = (lo > 0) | (hi < 0)
        with np.errstate(invalid='ignore', divide='ignore'):
            geo = np.sign(hi + lo) * np.sqrt(np.abs(lo) * np.abs(hi))
        ratio_big = same & (np.maximum(np.abs(lo), np.abs(hi)) > 4 * np.minimum(np.abs(lo), np.abs(hi)))
        mid_mu = np.where(ratio_big, geo, 0.5 * (lo + hi))
        # when one end is exactly 0 (the pole), step geometrically away from it
        zero_lo = (lo == 0); zero_hi = (hi == 0)
        mid_mu = np.where(zero_lo, hi * 1e-3 if it < 100 else 0.5 * hi, mid_mu)
        mid_mu = np.where(zero_hi, lo * 1e-3 if it < 100 else 0.5 * lo, mid_mu)
        gm = g(mid_mu)
        pos = gm > 0
        hi = np.where(pos, mid_mu, hi)
        lo = np.where(pos, lo, mid_mu)
        if np.all(np.abs(hi - lo) <= 2 * EPS * np.maximum(np.abs(lo), np.abs(hi))):
            break
    mu = 0.5 * (lo + hi)
    return org, mu, it


def secular_roots_bns(dd, z, rho, maxit=60):
    """Same contract as secular_roots; rational (Bunch-Nielsen-Sorensen two-pole) iteration with a bisection safeguard --
    the scheme of the CUDA kernel (one warp per root)."""
    k = len(dd)
    z2 = rho * z * z
    org = np.empty(k, dtype=int); mu = np.empty(k); its_max = 0
    for j in range(k):
        last = (j == k - 1)
        gap = (dd[j + 1] - dd[j]) if not last else z2.sum()
        # evaluate at the interval midpoint, differences against dd_j
        dl = dd - dd[j]
        x = 0.5 * gap
        r = 1.0 / (dl - x)
        t = z2 * r
        gmid = 1.0 + t.sum()
        if last or gmid >= 0:
            o = j; lo, hi = 0.0, (0.5 * gap if not last else gap); m = 0.5 * gap
        else:
            o = j + 1; lo, hi = -0.5 * gap, 0.0; m = -0.5 * gap
        dl = dd - dd[o]
        for it in range(maxit):
            r = 1.0 / (dl - m)
            t = z2 * r
            psi = t[:j + 1].sum(); phi = t[j + 1:].sum()
            dpsi = (t[:j + 1] * r[:j + 1]).sum(); dphi = (t[j + 1:] * r[j + 1:]).sum()
            g = 1.0 + psi + phi
            if g > 0: hi = min(hi, m)
            else: lo = max(lo, m)
            if abs(g) <= 8 * EPS * (1.0 + abs(psi) + abs(phi)) or (hi - lo) <= 2 * EPS * max(abs(lo), abs(hi)):
                break
            DL = dl[j] - m
            a = dpsi * DL * DL; c1 = psi - dpsi * DL
            if last:
                c = 1.0 + c1
                eta = (DL + a / c) if c > 0 else np.inf       # x = dL + a/c  -> eta = x - m
                cands = [eta]
            else:
                DR = dl[j + 1] - m
                b = dphi * DR * DR; c2 = phi - dphi * DR
                c = 1.0 + c1 + c2
                Bq = c * (DL + DR) + a + b
                Cq = DL * DR * g
                if c == 0:
                    cands = [Cq / Bq]
                else:
                    disc = Bq * Bq - 4 * c * Cq
                    sq = np.sqrt(max(disc, 0.0))
                    q = 0.5 * (Bq + (sq if Bq >= 0 else -sq))
                    cands = [q / c, (Cq / q) if q != 0 else np.inf]
            new = None
            for eta in cands:
                xm = m + eta
                if np.isfinite(xm) and lo < xm < hi:
                    new = xm; break
            if new is None:
                if lo == 0.0: new = 0.1 * hi
                elif hi == 0.0: new = 0.1 * lo
                else: new = 0.5 * (lo + hi)
            m = new
        org[j] = o; mu[j] = m; its_max = max(its_max, it)
    return org, mu, its_max


def merge(D1, Q1, D2, Q2, beta, stats):
    n1, n2 = len(D1), len(D2)
    N = n1 + n2
    rho = abs(beta)
    sgn = 1.0 if beta >= 0 else -1.0
    z = np.concatenate([Q1[-1, :], sgn * Q2[0, :]])
    D = np.concatenate([D1, D2])
    Q = np.zeros((N, N))
    Q[:n1, :n1] = Q1
    Q[n1:, n1:] = Q2
    # normalise z (||z||^2 = 2)
    z = z / np.sqrt(2.0)
    rho = 2.0 * rho
    perm = np.argsort(D, kind='stable')
    D = D[perm]; z = z[perm]; Q = Q[:, perm]
    tol = 8.0 * EPS * max(np.abs(D).max(), np.abs(z).max())
    defl = np.zeros(N, dtype=bool)
    if rho * np.abs(z).max() <= tol:
        stats['defl'] += N
        return D, Q
    small = rho * np.abs(z) <= tol
    defl |= small
    # close poles: sequential scan (as dlaed2)
    pj = -1
    nrot = 0
    for nj in range(N):
        if defl[nj]:
            continue
        if pj < 0:
            pj = nj
            continue
        s = z[pj]; c = z[nj]
        tau = np.hypot(c, s)
        t = D[nj] - D[pj]
        c /= tau; s = -s / tau
        if abs(t * c * s) <= tol:
            z[nj] = tau; z[pj] = 0.0
            qp = Q[:, pj].copy(); qn = Q[:, nj].copy()
            Q[:, pj] = c * qp + s * qn       # drot(Q(:,pj), Q(:,nj), c, s)
            Q[:, nj] = -s * qp + c * qn
            tt = D[pj] * c * c + D[nj] * s * s
            D[nj] = D[pj] * s * s + D[nj] * c * c
            D[pj] = tt
            defl[pj] = True
            nrot += 1
            pj = nj
        else:
            pj = nj
    nd = np.flatnonzero(~defl)
    df = np.flatnonzero(defl)
    k = len(nd)
    stats['defl'] += N - k; stats['rot'] += nrot; stats['tot'] += N
    dd = D[nd]; zz = z[nd]
    if k == 1:
        lam = np.array([dd[0] + rho * zz[0] ** 2])
        X = np.ones((1, 1))
    else:
        # D may have lost strict ordering among the non-deflated after rotations? (dlaed2 keeps order: D[nj] moves up, still <= next)
        assert np.all(np.diff(dd) >= 0), "non-deflated poles not sorted"
        org, mu, its = (secular_roots_bns if USE_BNS else secular_roots)(dd, zz, rho)
        stats['sec_its'] = max(stats['sec_its'], its)
        # lam_j - dd_i = (dd[org_j] - dd_i) + mu_j
        diff = (dd[org][None, :] - dd[:, None]) + mu[None, :]          # [i, j] = lam_j - dd_i
        lam = dd[org] + mu
        # Loewner / Gu-Eisenstat: zhat_i^2 = prod_j (lam_j - dd_i) / prod_{j != i} (dd_j - dd_i)  (/ rho)
        pd = dd[None, :] - dd[:, None]                                   # [i, j] = dd_j - dd_i
        np.fill_diagonal(pd, 1.0)
        # product in a stable interleaved order: ratio_j = (lam_j - dd_i)/(dd_j - dd_i) for j != i, times (lam_i - dd_i)
        ratio = diff / pd
        zhat2 = np.prod(ratio, axis=1) / rho
        zhat = np.sign(zz) * np.sqrt(np.abs(zhat2))
        X = zhat[:, None] / (-diff)                                     # x_ij = zhat_i / (dd_i - lam_j)
        X /= np.linalg.norm(X, axis=0)[None, :]
    Qn = Q[:, nd] @ X
    Dall = np.concatenate([lam, D[df]])
    Qall = np.concatenate([Qn, Q[:, df]], axis=1)
    o = np.argsort(Dall, kind='stable')
    return Dall[o], Qall[:, o]


def dc_tridiag(a, b, leaf=32, stats=None):
    """Eigen-decomposition of the symmetric tridiagonal (a, b) by Cuppen D&C, bottom-up over a fixed binary tree."""
    N = len(a)
    if stats is None:
        stats = {'defl': 0, 'rot': 0, 'tot': 0, 'sec_its': 0}
    # block boundaries: split until blocks <= leaf
    bounds = [0, N]
    while max(np.diff(bounds)) > leaf:
        nb = [0]
        for i in range(len(bounds) - 1):
            lo, hi = bounds[i], bounds[i + 1]
            if hi - lo > leaf:
                nb += [lo + (hi - lo) // 2, hi]
            else:
                nb += [hi]
        bounds = nb
    a = a.copy()
    # tear: subtract |beta| at both sides of each cut (all cuts, all levels, at once)
    cuts = bounds[1:-1]
    for c in cuts:
        a[c - 1] -= abs(b[c - 1]); a[c] -= abs(b[c - 1])
    blocks = []
    for i in range(len(bounds) - 1):
        lo, hi = bounds[i], bounds[i + 1]
        w, Q = leaf_eig(a[lo:hi], b[lo:hi - 1])
        blocks.append((lo, hi, w, Q))
    while len(blocks) > 1:
        nxt = []
        for i in range(0, len(blocks), 2):
            if i + 1 == len(blocks):
                nxt.append(blocks[i]); continue
            lo, mid, D1, Q1 = blocks[i]
            _, hi, D2, Q2 = blocks[i + 1]
            Dn, Qn = merge(D1, Q1, D2, Q2, b[mid - 1], stats)
            nxt.append((lo, hi, Dn, Qn))
        blocks = nxt
    return blocks[0][2], blocks[0][3], stats


def bidiag_svd_dc(d, e, leaf=32):
    n = len(d)
    off = np.empty(2 * n - 1)
    off[0::2] = d
    off[1::2] = e
    lam, Q, stats = dc_tridiag(np.zeros(2 * n), off, leaf)
    idx = np.argsort(-lam)[:n]
    s = lam[idx]
    q = Q[:, idx]
    V = q[0::2, :]; U = q[1::2, :]
    V = V / np.linalg.norm(V, axis=0); U = U / np.linalg.norm(U, axis=0)
    return U, s, V, stats


if __name__ == "__main__":
    sys.path.insert(0, '.')
    from oracle.kbdm_oracle import brain_sim, hankel_matrices, kbdm_oracle, compare_members, normalise
    from scipy.linalg import lapack, eig
    import time
    m = int(sys.argv[1]); sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
    c = brain_sim(2 * m, sigma, 0)
    U0, Up1, Up = hankel_matrices(c, m, 1)
    def house(x):
        alpha = x[0]; xn2 = np.sum(np.abs(x[1:]) ** 2)
        if xn2 == 0 and alpha.imag == 0: return np.zeros_like(x), 0.0, alpha
        beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn2), alpha.real)
        tau = (beta - alpha) / beta
        v = x / (alpha - beta); v[0] = 1
        return v, tau, beta
    A = U0.astype(complex).copy(); Qm = np.eye(m, dtype=complex); Pm = np.eye(m, dtype=complex)
    for k in range(m):
        v, tau, beta = house(A[k:, k].copy())
        if tau != 0:
            A[k:, k:] -= np.conj(tau) * np.outer(v, v.conj() @ A[k:, k:])
            Qm[:, k:] -= tau * np.outer(Qm[:, k:] @ v, v.conj())
        if k < m - 1:
            v, tau, beta = house(A[k, k + 1:].conj().copy())
            if tau != 0:
                A[k:, k + 1:] -= tau * np.outer(A[k:, k + 1:] @ v, v.conj())
                Pm[:, k + 1:] -= tau * np.outer(Pm[:, k + 1:] @ v, v.conj())
    d = np.diag(A).real.copy(); e = np.diag(A, 1).real.copy()
    print("bidiag: imag diag", np.abs(np.diag(A).imag).max(), "imag super", np.abs(np.diag(A, 1).imag).max(),
          "resid", np.abs(Qm @ (np.diag(d) + np.diag(e, 1)) @ Pm.conj().T - U0).max())
    t0 = time.time()
    Ub, s, Vb, stats = bidiag_svd_dc(d, e, leaf=32)
    print(f"dc time {time.time()-t0:.1f}s stats={stats}")
    B = np.diag(d) + np.diag(e, 1)
    sref = np.linalg.svd(B, compute_uv=False)
    print("sv rel err vs numpy (bidiag):", np.max(np.abs(s - sref) / sref), " abs/smax:", np.max(np.abs(s - sref)) / sref[0])
    print("orth U:", np.abs(Ub.T @ Ub - np.eye(m)).max(), " orth V:", np.abs(Vb.T @ Vb - np.eye(m)).max())
    print("resid |B - U S V^T|/|B|:", np.abs(B - (Ub * s) @ Vb.T).max() / s[0])
    L = Qm @ Ub; R = Pm @ Vb
    print("resid |U0 - L S R^H|/|U0|:", np.abs(U0 - (L * s) @ R.conj().T).max() / s[0])
    dsqi = 1 / np.sqrt(s)
    Ured = (dsqi[:, None] * (L.conj().T @ Up @ R)) * dsqi[None, :]
    mu, P = eig(Ured)
    Bm = R @ (dsqi[:, None] * P)
    Bn = normalise(Bm, U0)
    D = (c[:m] @ Bn) ** 2
    ll, info, mu_o, D_o = kbdm_oracle(c, 5e-4, m=m, return_mu=True)
    print("vs oracle: dmu, dD =", compare_members(mu, D, mu_o, D_o))
    print("sv vs oracle rel:", np.max(np.abs(s - info.singular_values) / info.singular_values))
