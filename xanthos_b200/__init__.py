"""
xanthos_b200 - the Xanthos per-month grid hot path (PET -> ABCD runoff -> MRTM routing and the ABCD
calibration objective) on NVIDIA B200, behind the reference's own Python surface.

    import xanthos_b200 as xanthos
    xanthos.run_model('pm_abcd_mrtm.ini')
    res = xanthos.Xanthos('pm_abcd_mrtm.ini').execute(args)      # Components with .PET .AET .Q .Sav ...

All arithmetic runs in libxanthos_b200.so (hand-written CUDA for sm_100a, C ABI in
include/xanthos_b200.h); there is no CPU fallback.  Importing the package does not load the library;
the first compute call does, and raises if it is missing.
"""

from .model import Xanthos, run_model
from .calibrate.calibrate_abcd import Calibrate

__all__ = ['Xanthos', 'run_model', 'Calibrate']
__version__ = '0.1.0'
