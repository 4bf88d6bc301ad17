"""
Load the *reference* hot-path modules.  TEST / BENCHMARK INFRASTRUCTURE ONLY.

Two locations:
  * /root/reference (build container only): the whole package; it cannot be imported as is here
    (configobj and matplotlib are not installed), so `configobj`/`matplotlib` are stubbed in
    sys.modules before `import xanthos`.  Used by oracle/validate_against_reference.py and
    tests/golden/make_golden.py.
  * oracle/_ref (staged by oracle/make_ref.sh, git-ignored, travels to the GPU box): the unmodified
    hot-path modules only, with empty package __init__ files - numpy / scipy / joblib are all they
    need.  Used by bench.py (`--impl reference`, cpu_baseline) through `load_staged()`.
"""

import os
import sys
from types import SimpleNamespace
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get('XANTHOS_REFERENCE_ROOT', '/root/reference')


STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'xanthos'))


def staged_available():
    return os.path.isfile(os.path.join(STAGED_ROOT, 'xanthos', 'routing', 'mrtm.py'))


def load_staged():
    """The unmodified hot-path modules staged under oracle/_ref (see oracle/make_ref.sh)."""
    if not staged_available():
        raise RuntimeError("no staged reference under {} - run oracle/make_ref.sh where /root/reference exists"
                           .format(STAGED_ROOT))
    if 'xanthos' in sys.modules and not getattr(sys.modules['xanthos'], '__file__', '').startswith(STAGED_ROOT):
        raise RuntimeError("another `xanthos` package is already imported in this process")
    if STAGED_ROOT not in sys.path:
        sys.path.insert(0, STAGED_ROOT)
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
                           hargreaves=hargreaves, gwam=gwam, root=STAGED_ROOT)


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
