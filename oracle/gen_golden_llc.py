"""Generate tests/golden/llc_*.npz by running the REAL reference's LLC-KBDM driver (read-only import of
/root/reference/llckbdm/llckbdm.py) in the build container.

Shims: ``numpy.complex = complex``; the un-vendored ``hdbscan`` package is replaced by a stub module whose
``HDBSCAN(min_samples=k)`` is ``sklearn.cluster.HDBSCAN(min_samples=k+1)`` -- hdbscan does not count the point itself among its
min_samples neighbours, sklearn does (the convention llckbdm_b200 documents and tests); every other parameter is the default of
both.  The reference's ragged ``np.array(clustered)`` (llckbdm.py:317) only works when all clusters have the same size, which
holds for the noiseless inputs used here (the reference's own test_llc_kbdm relies on the same fact).

    python oracle/gen_golden_llc.py
"""
import os
import sys
import types

import numpy as np

np.complex = complex
from sklearn.cluster import HDBSCAN as _SkHDBSCAN  # noqa: E402


class _HDBSCAN(_SkHDBSCAN):
    def __init__(self, min_samples=None, **kw):
        super().__init__(min_samples=None if min_samples is None else int(min_samples) + 1, copy=True, **kw)


stub = types.ModuleType("hdbscan")
stub.HDBSCAN = _HDBSCAN
sys.modules["hdbscan"] = stub
sys.path.insert(0, '/root/reference')
from llckbdm.llckbdm import llc_kbdm  # noqa: E402
from llckbdm import sig_gen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '..', 'tests', 'golden')


def brain_sim_ref(N, dwell=5e-4):
    import pandas as pd
    df = pd.read_csv('/root/reference/data/params_brain_sim_1_5T.csv',
                     names=['amplitude', 't2', 'frequency', 'phase']).sort_values(['frequency'])
    t = np.linspace(0, dwell * N, N, endpoint=False)
    return sig_gen.multi_fid(t, df.values)


if __name__ == "__main__":
    # (name, N, m_range, l) -- the first case is the reference's own test_llc_kbdm (llckbdm/_tests/test_llckbdm.py:37-57)
    for name, N, m_range, l in [("llc_clean_m250_l30", 2048, list(range(250, 260)), 30), ("llc_clean_m100_l40", 1024, list(range(100, 112)), 40)]:
        c = brain_sim_ref(N)
        r = llc_kbdm(c, 5e-4, m_range, l=l)
        np.savez(os.path.join(OUT, name + ".npz"), data=c, dwell=5e-4, m_range=np.array(m_range), l=l,
                 line_list=np.asarray(r.line_list), rmse=r.rmse, silhouette=np.asarray(r.silhouette))
        strong = r.line_list[r.line_list[:, 0] > 1e-3]
        print(name, "lines", len(r.line_list), "strong", len(strong), "rmse", r.rmse)
