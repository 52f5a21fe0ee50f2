"""Drop-in for reference llckbdm/min_rmse_kbdm.py: pick the candidate line list with minimum frequency-domain RMSE.
Candidates come from the batched GPU ``sample_kbdm`` unless given; the scoring loop (min_rmse_kbdm.py:33-41 over
metrics.py:7-17) runs on the device (``llck_rmse_batched``): fused with the solve when the candidates are the ensemble's own
line lists, one packed launch when they are given by the caller."""
import logging

import attr
import numpy as np

from .ensemble import score_candidates
from .sampling import sample_kbdm_scored
from .sig_gen import _validate_parameters

logger = logging.getLogger(__name__)


@attr.s
class MinRmseKbdmResult:
    """Same attrs record as reference min_rmse_kbdm.py:12-18."""
    line_list = attr.ib()
    min_rmse = attr.ib()
    min_index = attr.ib()
    samples = attr.ib()
    rmses_list = attr.ib()


def min_rmse_kbdm(data, dwell, m_range=None, l=None, samples=None):
    if samples is None:
        samples, _, rmses = sample_kbdm_scored(data=data, dwell=dwell, m_range=m_range, l=l, q=0, p=1,
                                               filter_invalid_features=True)          # min_rmse_kbdm.py:22-31
    else:
        for line_list in samples:                 # the reference's multi_fid validates every row (sig_gen.py:140-169)
            for row in line_list:
                _validate_parameters(*row)
        rmses = score_candidates(np.asarray(data).ravel(), dwell, samples, filter_rows=False)
    for i, rmse in enumerate(rmses):
        logger.debug('RMSE for sample #%d: %f', i, rmse)
    if not rmses:
        return None
    k = int(np.argmin(rmses))
    return MinRmseKbdmResult(line_list=samples[k], min_rmse=rmses[k], min_index=k, samples=samples, rmses_list=rmses)
