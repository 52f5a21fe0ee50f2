"""Drop-in for reference llckbdm/min_rmse_kbdm.py: pick the candidate line list with minimum
frequency-domain RMSE; candidates come from the batched GPU ``sample_kbdm`` unless given."""
import logging

import numpy as np

from .metrics import calculate_freq_domain_rmse
from .sampling import sample_kbdm

logger = logging.getLogger(__name__)


class MinRmseKbdmResult:
    __slots__ = ("line_list", "min_rmse", "min_index", "samples", "rmses_list")

    def __init__(self, line_list, min_rmse, min_index, samples, rmses_list):
        self.line_list = line_list
        self.min_rmse = min_rmse
        self.min_index = min_index
        self.samples = samples
        self.rmses_list = rmses_list


def min_rmse_kbdm(data, dwell, m_range=None, l=None, samples=None):
    if samples is None:
        samples, _ = sample_kbdm(data=data, dwell=dwell, m_range=m_range, l=l, q=0, p=1,
                                 filter_invalid_features=True)          # min_rmse_kbdm.py:22-31
    rmses = []
    for i, line_list in enumerate(samples):
        rmse = calculate_freq_domain_rmse(data=data, params_est=line_list, dwell=dwell) if len(line_list) > 0 else np.inf
        rmses.append(rmse)
        logger.debug('RMSE for sample #%d: %f', i, rmse)
    if not rmses:
        return None
    k = int(np.argmin(rmses))
    return MinRmseKbdmResult(line_list=samples[k], min_rmse=rmses[k], min_index=k, samples=samples, rmses_list=rmses)
