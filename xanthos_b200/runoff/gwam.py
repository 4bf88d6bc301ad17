"""
GWAM runoff on the B200 - drop-in for xanthos/runoff/gwam.py (the step-wise v1 runoff model).

`runoffgen(PET, P, settings, Sm, chstor, indexing=999)` keeps the reference's per-month signature and
return value `[PET, AET, Q, Sav]` (gwam.py:18-88).  `Components.simulation` does not loop over months:
`run_device` carries the soil moisture from month to month in registers, for the spin-up pass and the
simulation alike (components.py:358-366, configurations.py:106-123).
"""

import numpy as np

from .. import _cuda as C

LAKE = 999          # gwam.py:18 `indexing`: maximum soil moisture value that marks water bodies


def run_device(pet, precip, sm_max, sm_prev, n_months=None, spinup_months=0, want=('aet', 'q', 'sav')):
    """
    pet, precip: [ncell, nmonths] host arrays or Fields (precip keeps its NaNs); sm_max, sm_prev [ncell].
    `spinup_months` months from `sm_prev` that only carry the soil moisture over, then `n_months` from
    month 0.  Returns dict of Fields (aet, q, sav) and device vectors sm_after_spinup, sm_last.
    """
    torch = C.torch_cuda()
    e, p = C.as_field(pet), C.as_field(precip)
    if (e.ncell, e.ld) != (p.ncell, p.ld):
        raise C.ValidationException("gwam: PET and precipitation differ in shape")
    m = int(n_months) if n_months is not None else e.nmonths
    if m > e.nmonths or m > p.nmonths:
        raise C.ValidationException("gwam: n_months exceeds the forcing")
    smx, smp = C.dev_vector(sm_max), C.dev_vector(sm_prev)
    out = {k: C.Field.empty(e.ncell, m, e.ld) for k in want}
    spun = torch.empty(e.ncell, dtype=torch.float64, device='cuda')
    last = torch.empty(e.ncell, dtype=torch.float64, device='cuda')
    C.check(C.lib().xan_gwam_run(C.ptr(e.t), C.ptr(p.t), C.ptr(smx), C.ptr(smp), e.ncell, m, int(spinup_months), e.ld,
                                 C.ptr(out['aet'].t if 'aet' in out else None),
                                 C.ptr(out['q'].t if 'q' in out else None),
                                 C.ptr(out['sav'].t if 'sav' in out else None), C.ptr(spun), C.ptr(last),
                                 C.stream_ptr()))
    out['sm_after_spinup'], out['sm_last'] = spun, last
    return out


def runoffgen(PET, P, settings, Sm, chstor, indexing=LAKE):
    """One month, reference signature (gwam.py:18-88) -> [PET, AET, Q, Sav], each [ncell]."""
    if indexing != LAKE:
        raise C.ValidationException("gwam: only the reference's water-body marker 999 is supported")
    pet = np.asarray(PET, dtype=np.float64).reshape(-1)
    n = pet.shape[0]
    res = run_device(pet.reshape(n, 1), np.asarray(P, dtype=np.float64).reshape(n, 1), Sm, chstor, 1, 0)
    return [PET, res['aet'].to_host()[:, 0].copy(), res['q'].to_host()[:, 0].copy(), res['sav'].to_host()[:, 0].copy()]
