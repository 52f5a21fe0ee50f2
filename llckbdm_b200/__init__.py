"""llckbdm_b200 -- B200-native (sm_100a) drop-in for the KBDM / LLC-KBDM hot path of danilomendesdias/llckbdm.

Sub-modules mirror the reference layout (the reference's ``llckbdm/__init__.py`` is empty and users
import sub-modules by path): ``kbdm``, ``sampling``, ``min_rmse_kbdm``, ``llckbdm``, ``metrics``,
``sig_gen``.  The per-member KBDM solve runs on the GPU through the C ABI in ``include/llck.h``.
"""
__version__ = "0.1.0"
