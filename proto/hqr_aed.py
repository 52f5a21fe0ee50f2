"""Prototype: small-bulge multishift QR + aggressive early deflation (AED) -- counts sweeps/flops vs no AED."""
import numpy as np, sys, time
import scipy.linalg as sla
sys.path.insert(0, 'proto'); sys.path.insert(0, '.')
from hqr_multishift import givens, rot_rows, rot_cols, negligible, small_hqr, multishift_sweep, cabs1, EPS, SMALL

def trexc_up(T, V, ifst, ilst):
    """move T[ifst,ifst] up to position ilst by adjacent swaps (complex upper triangular)"""
    n = T.shape[0]
    for k in range(ifst - 1, ilst - 1, -1):
        t11 = T[k, k]; t22 = T[k + 1, k + 1]
        c, s = givens(T[k, k + 1], t22 - t11)
        # apply rotation G = [c s; -conj(s) c] to rows k,k+1 (cols k+2..) ; cols k,k+1 (rows 0..k-1) with G^H
        if k + 2 < n:
            rot_rows(T, k, c, s, k + 2, n)
        rot_cols(T, k, c, s, 0, k)
        T[k, k] = t22; T[k + 1, k + 1] = t11
        rot_cols(V, k, c, s, 0, V.shape[0])

def small_hess(T, V):
    """unblocked Householder Hessenberg of small T, accumulate V <- V Q"""
    n = T.shape[0]
    for k in range(n - 2):
        x = T[k + 1:, k].copy()
        alpha = x[0]; xn2 = np.sum(np.abs(x[1:]) ** 2)
        if xn2 == 0 and alpha.imag == 0: continue
        beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn2), alpha.real)
        tau = (beta - alpha) / beta
        v = x / (alpha - beta); v[0] = 1
        T[k + 1:, k:] -= np.conj(tau) * np.outer(v, v.conj() @ T[k + 1:, k:])
        T[:, k + 1:] -= tau * np.outer(T[:, k + 1:] @ v, v.conj())
        V[:, k + 1:] -= tau * np.outer(V[:, k + 1:] @ v, v.conj())
        T[k + 2:, k] = 0

def aed(H, Z, ilo, ihi, nw, stats):
    """returns (nd deflated, shifts)"""
    n = H.shape[0]
    nw = min(nw, ihi - ilo + 1)
    kwtop = ihi - nw + 1
    s = H[kwtop, kwtop - 1] if kwtop > ilo else 0.0
    T = H[kwtop:ihi + 1, kwtop:ihi + 1].copy()
    V = np.eye(nw, dtype=complex)
    r = small_hqr(T, V); assert r >= 0
    ns = nw; ilst = 0
    for knt in range(nw):
        foo = cabs1(T[ns - 1, ns - 1])
        if foo == 0: foo = cabs1(s)
        if cabs1(s) * cabs1(V[0, ns - 1]) <= max(SMALL, EPS * foo):
            ns -= 1
        else:
            trexc_up(T, V, ns - 1, ilst)
            ilst += 1
        if ilst >= ns: break
    if ns == 0: s = 0.0
    shifts = np.diag(T)[:ns].copy()
    nd = nw - ns
    if nd > 0 or s == 0:
        if ns > 1 and s != 0:
            # reflect spike back
            w = s * np.conj(V[0, :ns])          # spike (column vector entries)
            alpha = w[0]; xn2 = np.sum(np.abs(w[1:]) ** 2)
            if not (xn2 == 0 and alpha.imag == 0):
                beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn2), alpha.real)
                tau = (beta - alpha) / beta
                v = w / (alpha - beta); v[0] = 1
                # apply P = I - tau v v^H : T <- P^H T P on leading ns
                T[:ns, :] -= np.conj(tau) * np.outer(v, v.conj() @ T[:ns, :])
                T[:, :ns] -= tau * np.outer(T[:, :ns] @ v, v.conj())
                V[:, :ns] -= tau * np.outer(V[:, :ns] @ v, v.conj())
            Tn = T[:ns, :ns]
            # Hessenberg-reduce leading ns block, apply also to T[:ns, ns:] and V[:, :ns]
            Vn = np.eye(ns, dtype=complex)
            small_hess(Tn, Vn)
            T[:ns, ns:] = Vn.conj().T @ T[:ns, ns:]
            V[:, :ns] = V[:, :ns] @ Vn
        if kwtop > ilo:
            H[kwtop, kwtop - 1] = s * np.conj(V[0, 0])
        H[kwtop:ihi + 1, kwtop:ihi + 1] = np.triu(T, -1)
        if ihi + 1 < n:
            H[kwtop:ihi + 1, ihi + 1:] = V.conj().T @ H[kwtop:ihi + 1, ihi + 1:]
        if kwtop > 0:
            H[:kwtop, kwtop:ihi + 1] = H[:kwtop, kwtop:ihi + 1] @ V
        Z[:, kwtop:ihi + 1] = Z[:, kwtop:ihi + 1] @ V
        stats['flops'] += 8.0 * nw * nw * ((n - ihi - 1) + kwtop + n)
        stats['aed_applied'] += 1
    stats['aed'] += 1
    return nd, shifts

def hqr_aed(H, Z, nb=16, w=64, nw=32, nibble=0.14):
    n = H.shape[0]
    ihi = n - 1
    stats = dict(windows=0, flops=0.0, sweeps=0, small=0, aed=0, aed_applied=0)
    its = 0
    while ihi >= 0:
        ilo = ihi
        while ilo > 0 and not negligible(H, ilo):
            ilo -= 1
        if ilo > 0: H[ilo, ilo - 1] = 0
        if ilo == ihi:
            ihi -= 1; its = 0; continue
        size = ihi - ilo + 1
        if size <= w:
            Hw = H[ilo:ihi + 1, ilo:ihi + 1].copy(); Ww = np.eye(size, dtype=complex)
            r = small_hqr(Hw, Ww); assert r >= 0
            H[ilo:ihi + 1, ilo:ihi + 1] = Hw
            if ihi + 1 < n: H[ilo:ihi + 1, ihi + 1:n] = Ww.conj().T @ H[ilo:ihi + 1, ihi + 1:n]
            if ilo > 0: H[0:ilo, ilo:ihi + 1] = H[0:ilo, ilo:ihi + 1] @ Ww
            Z[:, ilo:ihi + 1] = Z[:, ilo:ihi + 1] @ Ww
            stats['small'] += 1; stats['flops'] += 8.0 * size * size * ((n - ihi - 1) + ilo + n)
            ihi = ilo - 1; its = 0; continue
        its += 1
        assert its < 100
        nd, shifts = aed(H, Z, ilo, ihi, nw, stats)
        ihi -= nd
        if nd > 0: its = 0
        if nd > nibble * nw or ihi - ilo + 1 <= w:
            continue     # good deflation: try AED again without a sweep
        ns = len(shifts)
        if ns < 2 or its % 6 == 0:
            shifts = np.array([H[ihi - i, ihi - i] + 0.75 * abs(H[ihi - i, ihi - i - 1]) for i in range(nb)])
        elif ns > nb:
            shifts = shifts[-nb:] if False else shifts[:nb]
        multishift_sweep(H, Z, ilo, ihi, list(shifts), w, stats)
        stats['sweeps'] += 1
    return stats

if __name__ == '__main__':
    n = int(sys.argv[1]); nb = int(sys.argv[2]); nw = int(sys.argv[3])
    from oracle.kbdm_oracle import brain_sim, hankel_matrices, reduce_gep
    c = brain_sim(2048, 1e-3, 0)
    U0, Up1, Up = hankel_matrices(c, n, 1)
    A, _, _, _ = reduce_gep(Up1, Up, n)
    H, Q = sla.hessenberg(A, calc_q=True)
    Z = Q.copy()
    t0 = time.time()
    st = hqr_aed(H, Z, nb=nb, nw=nw)
    print("time", time.time() - t0, st, "flops/n^3", st['flops'] / n ** 3)
    print("tril", np.abs(np.tril(H, -1)).max(), "resid", np.abs(Z @ H @ Z.conj().T - A).max() / np.abs(A).max(), "orth", np.abs(Z.conj().T @ Z - np.eye(n)).max())
    ev = np.diag(H); ref = np.linalg.eigvals(A)
    from scipy.optimize import linear_sum_assignment
    cost = np.abs(ev[:, None] - ref[None, :]); r, cidx = linear_sum_assignment(cost)
    print("max eig err", np.max(np.abs(ev[r] - ref[cidx]) / np.abs(ref[cidx])))
