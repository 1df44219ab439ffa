"""B200-native thermal hot path of AdditiveFOAM (assembly, moving heat source, solidification
corrector loop, PCG/DIC and PBiCGStab/DILU on LDU addressing) behind a C ABI.

The product is ``csrc/`` (CUDA for sm_100a + the C ABI in ``include/additivefoam_b200.h``);
this package is the thin Python host used by the tests and ``bench.py``.  There is no CPU
fallback: importing :mod:`additivefoam_b200.lib` fails loudly when the CUDA library is missing.
"""
from . import abi, cases  # noqa: F401
