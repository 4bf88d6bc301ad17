"""
Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE
(/root/reference, build container only):

    python tests/golden/make_golden.py

Each fixture is a compressed .npz holding every input array of a parity case
(`in_*`) and the outputs of the reference's own functions (`ref_*`):
run_pmpet, hargreaves_samani.execute, thornthwaite.execute, abcd_execute
(jobs=1, jobs=-1, no-snow), downstream / upstream / upstream_genmatrix,
streamrouting driven like Components.calculate_routing, and objective_kge; case_c holds the
step-wise path (hargreaves.calculate_pet, calc_sinusoidal_factor, gwam.runoffgen with its spin-up pass);
case_d the post-processing scans (DroughtStats.calculate_thresholds / droughtstats, Aggregation_Map, the
accessible-water chain).
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402
from oracle.validate_against_reference import (build_case, run_reference, build_stepwise_case,  # noqa: E402
                                               run_reference_stepwise, build_postproc_case, run_reference_postproc)

CASES = {
    # 3 years around the leap year 2000, spin-up = whole period
    'case_a': dict(nrow=24, ncol=48, ncell=300, n_basins=6, start_yr=1999, end_yr=2001, seed=11,
                   spinup=36, routing_spinup=4),
    # ends in 2100: the mod-4 and Gregorian leap rules disagree
    'case_b': dict(nrow=18, ncol=36, ncell=150, n_basins=4, start_yr=2098, end_yr=2100, seed=23,
                   spinup=25, routing_spinup=36),
}


def main():
    ref = ref_loader.load()
    for name, kw in CASES.items():
        case = build_case(**kw)
        out = run_reference(case, ref)
        blob = {'in_' + k: np.asarray(v) for k, v in case.items()}
        blob.update({'ref_' + k: np.asarray(v) for k, v in out.items()})
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **blob)
        print(name, '->', path, '{:.2f} MB'.format(os.path.getsize(path) / 1e6))
    # step-wise (v1) path: Hargreaves PET + GWAM runoff with its spin-up pass; 2003-2005 (leap year 2004)
    case = build_stepwise_case(nrow=24, ncol=48, ncell=300, n_basins=6, start_yr=2003, end_yr=2005, seed=31, spinup=14)
    out = run_reference_stepwise(case, ref)
    blob = {'in_' + k: np.asarray(v) for k, v in case.items()}
    blob.update({'ref_' + k: np.asarray(v) for k, v in out.items()})
    path = os.path.join(HERE, 'case_c.npz')
    np.savez_compressed(path, **blob)
    print('case_c', '->', path, '{:.2f} MB'.format(os.path.getsize(path) / 1e6))
    # post-processing scans (drought statistics, Aggregation_Map, accessible water), 1971-2001
    case = build_postproc_case()
    out = run_reference_postproc(case)
    blob = {'in_' + k: np.asarray(v) for k, v in case.items()}
    blob.update({'ref_' + k: np.asarray(v) for k, v in out.items()})
    path = os.path.join(HERE, 'case_d.npz')
    np.savez_compressed(path, **blob)
    print('case_d', '->', path, '{:.2f} MB'.format(os.path.getsize(path) / 1e6))


if __name__ == '__main__':
    main()
