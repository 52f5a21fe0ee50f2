import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import _native
from oracle.kbdm_oracle import brain_sim, hankel_matrices
lib = _native.load()
dev = torch.device("cuda:0")
for m in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "3,31,32,33,64,100,257").split(",")]:
    ld = lib.llck_leading_dim(m)
    if m <= 64:
        rng = np.random.default_rng(m); A = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
    else:
        A, _, _ = hankel_matrices(brain_sim(2 * m + 8, 1e-3, 0), m, 1)
    Ap = np.zeros((ld, ld), dtype=complex); Ap[:m, :m] = A
    Ad = torch.from_numpy(np.ascontiguousarray(Ap.T).view(np.float64)).to(dev)      # column-major
    Qd = torch.zeros((ld, ld, 2), dtype=torch.float64, device=dev); Pd = torch.zeros_like(Qd)
    dd = torch.zeros(ld, dtype=torch.float64, device=dev); ed = torch.zeros(ld, dtype=torch.float64, device=dev)
    rc = lib.llck_bidiag_test(Ad.data_ptr(), m, ld, dd.data_ptr(), ed.data_ptr(), Qd.data_ptr(), Pd.data_ptr(), None)
    assert rc == 0, rc
    Q = Qd.cpu().numpy().view(np.complex128)[..., 0].T[:m, :m]; P = Pd.cpu().numpy().view(np.complex128)[..., 0].T[:m, :m]
    d = dd.cpu().numpy()[:m]; e = ed.cpu().numpy()[:m - 1]
    B = np.diag(d) + np.diag(e, 1)
    s1 = np.linalg.svd(B, compute_uv=False); s0 = np.linalg.svd(A, compute_uv=False)
    print(f"m={m}: resid={np.abs(Q @ B @ P.conj().T - A).max() / np.abs(A).max():.2e} Qorth={np.abs(Q.conj().T @ Q - np.eye(m)).max():.2e} "
          f"Porth={np.abs(P.conj().T @ P - np.eye(m)).max():.2e} sv rel={np.max(np.abs(s1 - s0) / s0):.2e}")
