"""GPU stage-by-stage diagnostic: runs llck_kbdm_batched with LLCK_FLAG_DEBUG_KEEP and checks every
intermediate of every member against numpy applied to the PREVIOUS stage's GPU output.

    python tools/stage_check.py [m1,m2,...] [l or -1] [p] [q]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch  # noqa: E402

from llckbdm_b200 import _native, ensemble  # noqa: E402
from oracle.kbdm_oracle import brain_sim, hankel_matrices, kbdm_oracle, compare_members  # noqa: E402

NAMES = ["X", "V", "Rs", "Lt", "T1", "Ured", "Hhess", "Qhess", "T", "Z", "Xev", "P", "B", "W"]


def get_mats(ws, lib, batch, ld):
    out = {}
    buf = ws.cpu().numpy()
    for i, nm in enumerate(NAMES):
        off = lib.llck_debug_offset(batch, ld, i)
        nbytes = batch * ld * ld * 16
        a = buf[off:off + nbytes].view(np.complex128).reshape(batch, ld, ld)
        out[nm] = a.transpose(0, 2, 1)   # column-major -> [b, row, col]
    return out


def rel(a, b):
    d = np.abs(b).max()
    return np.abs(a - b).max() / (d if d > 0 else 1.0)


def main():
    ms = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "16,64,100").split(",")]
    lsel = int(sys.argv[2]) if len(sys.argv) > 2 else -1
    p = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    q = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
    sigma = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-3
    dwell = 5e-4
    lib = _native.load()
    sigs = [brain_sim(max(2048, 2 * max(ms) + 8), sigma, seed=i) for i in range(len(ms))]
    ls = [m if lsel < 0 else min(lsel, m) for m in ms]
    flat, offs, lens = ensemble.flatten_signals(sigs, len(ms))
    dev = torch.device("cuda:0")
    sig_dev = torch.from_numpy(flat.view(np.float64)).to(dev).view(torch.complex128)
    torch.cuda.synchronize()
    t0 = time.time()
    ws_bytes = lib.llck_workspace_bytes(len(ms), lib.llck_leading_dim(max(ms)), _native.FLAG_DEBUG_KEEP)
    ws_dbg = torch.empty(ws_bytes, dtype=torch.uint8, device=sig_dev.device)
    r = ensemble.solve_device(sig_dev, offs, ms, ls, p, q, dwell, flags=_native.FLAG_DEBUG_KEEP, workspace=ws_dbg, sig_len=lens)
    torch.cuda.synchronize()
    print(f"solve time {time.time() - t0:.3f}s info={r['info']} status={r['status'].cpu().tolist()} n_valid={r['n_valid'].cpu().tolist()}")
    ld = r["ld"]
    M = get_mats(ws_dbg, lib, len(ms), ld)
    sv = r["sing_vals"].cpu().numpy()
    ll = r["line_lists"].cpu().numpy()
    mu = r["mu"].cpu().numpy()
    D = r["D"].cpu().numpy()
    ok_all = True
    for b, (m, l) in enumerate(zip(ms, ls)):
        c = sigs[b]
        U0, Up1, Up = hankel_matrices(c, m, p)
        mp = ((m + 63) // 64) * 64
        X = M["X"][b][:m, :mp]; V = M["V"][b][:m, :mp]
        s_ref = np.linalg.svd(Up1, compute_uv=False)
        s_gpu = sv[b, :m]
        e_sv = np.max(np.abs(s_gpu - s_ref) / s_ref)
        e_xv = rel(Up1 @ V, X)
        G = X.conj().T @ X
        dn = np.sqrt(np.abs(np.diag(G))); dn[dn == 0] = 1
        e_orth = np.abs((G - np.diag(np.diag(G))) / np.outer(dn, dn)).max()
        e_vorth = np.abs(V[:, :m].conj().T @ V[:, :m] - np.eye(m)).max() if mp == m else np.abs(V @ V.conj().T - np.eye(m)).max()
        Rs = M["Rs"][b][:m, :l]; Lt = M["Lt"][b][:m, :l]
        g = s_gpu[:l] if q == 0 else s_gpu[:l] + q * q / s_gpu[:l]
        # Rs/Lt consistency: Lt * g * Rs^H ~ best rank-l approx; check Lt^H Up1 Rs = diag(1/g * s)?  simpler: Up1 @ Rs = Lt * s
        e_rl = rel(Up1 @ Rs, Lt * s_gpu[:l])
        T1 = M["T1"][b][:m, :l]
        e_t1 = rel(T1, Up @ Rs)
        Ured = M["Ured"][b][:l, :l]
        e_ur = rel(Ured, Lt.conj().T @ T1)
        Hh = M["Hhess"][b][:l, :l]; Qh = M["Qhess"][b][:l, :l]
        e_hq = rel(Qh @ Hh @ Qh.conj().T, Ured)
        e_qorth = np.abs(Qh.conj().T @ Qh - np.eye(l)).max()
        e_hess = np.abs(np.tril(Hh, -2)).max() if l > 2 else 0.0
        T = M["T"][b][:l, :l]; Z = M["Z"][b][:l, :l]
        e_schur = rel(Z @ T @ Z.conj().T, Ured)
        e_zorth = np.abs(Z.conj().T @ Z - np.eye(l)).max()
        e_tri = np.abs(np.tril(T, -1)).max()
        Xev = M["Xev"][b][:l, :l]
        lam = np.diag(T)
        e_trevc = np.abs(T @ Xev - Xev * lam[None, :]).max() / max(np.abs(T).max(), 1e-300)
        P = M["P"][b][:l, :l]
        e_p = rel(P, Z @ Xev)
        e_eig = np.max(np.linalg.norm(Ured @ P - P * lam[None, :], axis=0) / (np.linalg.norm(P, axis=0) * np.linalg.norm(Ured, 2)))
        Bm = M["B"][b][:m, :l]
        e_b = rel(Bm, Rs @ P)
        W = M["W"][b][:m, :l]
        e_w = rel(W, U0 @ Bm)
        Nk = (Bm * W).sum(0)
        Dk = W[0] ** 2 / Nk
        e_d = np.max(np.abs(D[b, :l] - Dk) / np.abs(Dk))
        e_mu = np.abs(mu[b, :l] - lam).max()
        # end-to-end vs oracle
        ll_o, info_o, mu_o, D_o = kbdm_oracle(c, dwell, m=m, p=p, l=l, q=q, return_mu=True)
        dmu, dD = compare_members(mu[b, :l], D[b, :l], mu_o, D_o)
        print(f"[m={m} l={l}] sv={e_sv:.1e} XV={e_xv:.1e} orthX={e_orth:.1e} orthV={e_vorth:.1e} RL={e_rl:.1e} T1={e_t1:.1e} Ured={e_ur:.1e} | "
              f"hessQHQ={e_hq:.1e} Qorth={e_qorth:.1e} hesslow={e_hess:.1e} | schur={e_schur:.1e} Zorth={e_zorth:.1e} tri={e_tri:.1e} "
              f"trevc={e_trevc:.1e} P={e_p:.1e} eigres={e_eig:.1e} | B={e_b:.1e} W={e_w:.1e} D={e_d:.1e} mu={e_mu:.1e} || ORACLE dmu={dmu:.2e} dD={dD:.2e}")
        if not (dmu < 1e-8 and dD < 1e-8 and e_sv < 1e-8):
            ok_all = False
    print("ALL_OK" if ok_all else "SOME_FAILED")


if __name__ == "__main__":
    main()
