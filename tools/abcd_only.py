"""ABCD passes alone on the bench world (for ncu / A-B timing): python tools/abcd_only.py [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xanthos_b200 import _cuda as C
from xanthos_b200.runoff import abcd as abcd_mod
world, pm, ab, end_yr = bench.build_inputs()
rows = abcd_mod._basin_rows(world.n_basins, world.basin_ids, world.n_basins)
plan = abcd_mod.basin_plan(rows, world.n_basins)
d_pars = torch.from_numpy(ab['pars']).cuda()
pet = C.Field.from_host(np.abs(np.random.default_rng(0).normal(80, 30, ab['precip'].shape)))
pr, tm = C.Field.from_host(ab['precip']), C.Field.from_host(ab['tmin'])
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(3):
    abcd_mod.run_device(plan, d_pars, pet, pr, tm, 360, 360)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    abcd_mod.run_device(plan, d_pars, pet, pr, tm, 360, 360)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print('abcd spin-up + re-init + simulation: %.4f ms  -> %.0f GB/s algorithmic (72 B per cell-month)' % (ms, 72 * 67420 * 360 / ms / 1e6))
