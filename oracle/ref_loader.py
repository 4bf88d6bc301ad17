"""
Load the *reference* hot-path modules from /root/reference (build container only).

The reference package cannot be imported as a whole here (configobj and
matplotlib are not installed), so `configobj`/`matplotlib` are stubbed in
sys.modules before `import xanthos`.  Used only by
oracle/validate_against_reference.py and tests/golden/make_golden.py; the GPU
box has no /root/reference and never calls this.
"""

import os
import sys
from types import SimpleNamespace
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get('XANTHOS_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'xanthos'))


def load():
    """Return a namespace with the reference modules of the hot path."""
    if not available():
        raise RuntimeError("reference tree not found at {}".format(REFERENCE_ROOT))
    for name in ('configobj', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.dates'):
        sys.modules.setdefault(name, MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import xanthos.pet.penman_monteith as pm
    import xanthos.pet.hargreaves_samani as hs
    import xanthos.pet.thornthwaite as tw
    import xanthos.runoff.abcd as abcd
    import xanthos.routing.mrtm as mrtm
    import xanthos.calibrate.calibrate_abcd as cal
    import xanthos.utils.general as general
    import xanthos.pet.hargreaves as hargreaves
    import xanthos.runoff.gwam as gwam
    return SimpleNamespace(pm=pm, hs=hs, tw=tw, abcd=abcd, mrtm=mrtm, cal=cal, general=general,
                           hargreaves=hargreaves, gwam=gwam)
