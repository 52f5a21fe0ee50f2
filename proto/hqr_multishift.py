"""Prototype of the windowed small-bulge multishift QR (complex, single-shift bulges, spacing 3, lockstep chase).

Mirrors the structure of the CUDA kernel: window in 'shared memory' (Hw, Ww), lockstep steps with all bulges moving
at once, then strip updates by dense W.
"""
import numpy as np, sys, time
import scipy.linalg as sla

EPS = np.finfo(float).eps
SMALL = np.finfo(float).tiny / EPS

def cabs1(z):
    return abs(z.real) + abs(z.imag)

def givens(a, b):
    """c real, s complex with [c s; -conj(s) c] [a; b] = [r; 0]"""
    if b == 0:
        return 1.0, 0.0 + 0j
    if a == 0:
        return 0.0, np.conj(b) / abs(b)
    na = abs(a); nrm = np.hypot(na, abs(b))
    c = na / nrm
    s = (a / na) * np.conj(b) / nrm
    return c, s

def rot_rows(M, r, c, s, c0, c1):
    x = M[r, c0:c1].copy(); y = M[r + 1, c0:c1].copy()
    M[r, c0:c1] = c * x + s * y
    M[r + 1, c0:c1] = -np.conj(s) * x + c * y

def rot_cols(M, k, c, s, r0, r1):
    x = M[r0:r1, k].copy(); y = M[r0:r1, k + 1].copy()
    M[r0:r1, k] = c * x + np.conj(s) * y
    M[r0:r1, k + 1] = -s * x + c * y

def negligible(H, k):
    """is H[k,k-1] negligible (k>=1)"""
    h = cabs1(H[k, k - 1])
    if h <= SMALL:
        return True
    tst = cabs1(H[k - 1, k - 1]) + cabs1(H[k, k])
    if tst == 0:
        tst = np.abs(H).sum()  # crude fallback
    return h <= EPS * tst

def small_hqr(Hs, W=None, maxit=30):
    """single-shift QR on small upper Hessenberg Hs (in place -> upper triangular). W (if given) accumulates: Hs_in = W T W^H.
    returns number of qr sweeps or -1 on failure"""
    n = Hs.shape[0]
    ihi = n - 1
    its = 0; total = 0
    while ihi >= 0:
        # find ilo
        ilo = ihi
        while ilo > 0 and not negligible(Hs, ilo):
            ilo -= 1
        if ilo > 0:
            Hs[ilo, ilo - 1] = 0
        if ilo == ihi:
            ihi -= 1; its = 0
            continue
        its += 1; total += 1
        if its > maxit * 10:
            return -1
        # shift: wilkinson (eigenvalue of trailing 2x2 closer to h_ii)
        if its % 10 == 0:
            sh = Hs[ihi, ihi] + 0.75 * abs(Hs[ihi, ihi - 1].real)   # exceptional
        else:
            a = Hs[ihi - 1, ihi - 1]; b = Hs[ihi - 1, ihi]; cc = Hs[ihi, ihi - 1]; d = Hs[ihi, ihi]
            tr2 = 0.5 * (a + d); det = a * d - b * cc
            disc = np.sqrt(tr2 * tr2 - det + 0j)
            e1 = tr2 + disc; e2 = tr2 - disc
            sh = e1 if abs(e1 - d) < abs(e2 - d) else e2
        # chase
        x = Hs[ilo, ilo] - sh; y = Hs[ilo + 1, ilo]
        for k in range(ilo, ihi):
            if k > ilo:
                x = Hs[k, k - 1]; y = Hs[k + 1, k - 1]
            c, s = givens(x, y)
            rot_rows(Hs, k, c, s, max(k - 1, ilo) if k > ilo else k, n)
            if k > ilo:
                Hs[k + 1, k - 1] = 0
            rot_cols(Hs, k, c, s, 0, min(k + 3, ihi + 1))
            if W is not None:
                rot_cols(W, k, c, s, 0, W.shape[0])
    return total

def multishift_sweep(H, Z, ilo, ihi, shifts, w=64, stats=None, SP=2):
    """one small-bulge multishift sweep on active block [ilo, ihi] (inclusive) of H (n x n), accumulating into Z."""
    n = H.shape[0]
    nb = len(shifts)
    t = 0
    while True:
        p_last = ilo - 1 - SP * (nb - 1) + t      # position of last-introduced bulge
        if p_last > ihi - 2:
            break
        p_top = max(ilo - 1, p_last)
        p0 = ilo - 1 + t
        ws = max(ilo, p_top)
        we = min(ws + w, ihi + 1)
        if we == ihi + 1:
            T = (ihi - 2) - p_last + 1
        else:
            T = we - 3 - p0
        assert T >= 1
        ww = we - ws
        Hw = H[ws:we, ws:we].copy()
        Ww = np.eye(ww, dtype=complex)
        for step in range(T):
            tt = t + step
            # phase 0: compute rotations for all active bulges
            rots = []
            for i in range(nb):
                p = ilo - 1 - SP * i + tt
                if p < ilo - 1 or p > ihi - 2:
                    continue
                if p == ilo - 1:
                    x = Hw[ilo - ws, ilo - ws] - shifts[i]; y = Hw[ilo + 1 - ws, ilo - ws]
                else:
                    x = Hw[p + 1 - ws, p - ws]; y = Hw[p + 2 - ws, p - ws]
                c, s = givens(x, y)
                rots.append((p, c, s))
            # phase A: row rotations
            for (p, c, s) in rots:
                r = p + 1 - ws
                c0 = max(p - ws, 0)
                rot_rows(Hw, r, c, s, c0, ww)
                if p >= ilo:
                    Hw[r + 1, p - ws] = 0
            # phase B: col rotations on Hw and Ww
            for (p, c, s) in rots:
                k = p + 1 - ws
                r1 = min(p + 3, ihi) - ws + 1
                rot_cols(Hw, k, c, s, 0, r1)
                rot_cols(Ww, k, c, s, 0, ww)
        H[ws:we, ws:we] = Hw
        # strips
        if we < n:
            H[ws:we, we:n] = Ww.conj().T @ H[ws:we, we:n]
        if ws > 0:
            H[0:ws, ws:we] = H[0:ws, ws:we] @ Ww
        Z[:, ws:we] = Z[:, ws:we] @ Ww
        if stats is not None:
            stats['windows'] += 1; stats['flops'] += 8.0 * ww * ww * ((n - we) + ws + n)
        t += T

def hqr_multishift(H, Z, nb=8, w=64, verbose=False):
    n = H.shape[0]
    ihi = n - 1
    stats = dict(windows=0, flops=0.0, sweeps=0, small=0)
    its = 0
    while ihi >= 0:
        ilo = ihi
        while ilo > 0 and not negligible(H, ilo):
            ilo -= 1
        if ilo > 0:
            H[ilo, ilo - 1] = 0
        if ilo == ihi:
            ihi -= 1; its = 0
            continue
        size = ihi - ilo + 1
        if size <= w:
            # whole active block fits in a window: finish it in 'shared memory'
            Hw = H[ilo:ihi + 1, ilo:ihi + 1].copy()
            Ww = np.eye(size, dtype=complex)
            r = small_hqr(Hw, Ww)
            assert r >= 0
            H[ilo:ihi + 1, ilo:ihi + 1] = Hw
            if ihi + 1 < n:
                H[ilo:ihi + 1, ihi + 1:n] = Ww.conj().T @ H[ilo:ihi + 1, ihi + 1:n]
            if ilo > 0:
                H[0:ilo, ilo:ihi + 1] = H[0:ilo, ilo:ihi + 1] @ Ww
            Z[:, ilo:ihi + 1] = Z[:, ilo:ihi + 1] @ Ww
            stats['small'] += 1; stats['flops'] += 8.0 * size * size * ((n - ihi - 1) + ilo + n)
            ihi = ilo - 1; its = 0
            continue
        its += 1
        if its > 60:
            raise RuntimeError("no convergence")
        # shifts = eigenvalues of trailing nb x nb
        Hs = H[ihi - nb + 1:ihi + 1, ihi - nb + 1:ihi + 1].copy()
        if its % 6 == 0:
            shifts = np.array([H[ihi - i, ihi - i] + 0.75 * abs(H[ihi - i, ihi - i - 1]) for i in range(nb)])
        else:
            r = small_hqr(Hs)
            assert r >= 0
            shifts = np.diag(Hs).copy()
        multishift_sweep(H, Z, ilo, ihi, shifts, w, stats)
        stats['sweeps'] += 1
        if verbose:
            print(f"sweep {stats['sweeps']}: active [{ilo},{ihi}] sub={abs(H[ihi, ihi-1]):.2e}")
    return stats

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    kind = sys.argv[3] if len(sys.argv) > 3 else 'kbdm'
    if kind == 'rand':
        rng = np.random.default_rng(0)
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    else:
        sys.path.insert(0, '.')
        from oracle.kbdm_oracle import brain_sim, hankel_matrices, reduce_gep
        c = brain_sim(2048, 1e-3, 0)
        U0, Up1, Up = hankel_matrices(c, n, 1)
        A, _, _, _ = reduce_gep(Up1, Up, n)
    H, Q = sla.hessenberg(A, calc_q=True)
    H0 = H.copy()
    Z = Q.copy()
    t0 = time.time()
    st = hqr_multishift(H, Z, nb=nb)
    print("time", time.time() - t0, st, "flops/n^3", st['flops'] / n ** 3)
    print("tril", np.abs(np.tril(H, -1)).max(), "resid", np.abs(Z @ H @ Z.conj().T - A).max() / np.abs(A).max(), "orth", np.abs(Z.conj().T @ Z - np.eye(n)).max())
    ev = np.diag(H); ref = np.linalg.eigvals(A)
    from scipy.optimize import linear_sum_assignment
    cost = np.abs(ev[:, None] - ref[None, :]); r, cidx = linear_sum_assignment(cost)
    print("max eig err", np.max(np.abs(ev[r] - ref[cidx]) / np.abs(ref[cidx])))
