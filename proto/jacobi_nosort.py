import numpy as np, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'proto')
from oracle.kbdm_oracle import brain_sim, hankel_matrices
from jacobi_svd4 import run
m = int(sys.argv[1]); b = 32
c = brain_sim(2 * m, 1e-3, 0)
U0, _, _ = hankel_matrices(c, m, 1)
for sort in (True, False):
    ns, ti, hist = run(U0, b, 1, conv=1e-6, sort=sort)
    print(f"sort={sort} outer sweeps={ns} hist=" + " ".join(f"{h:.1e}" for h in hist), flush=True)
