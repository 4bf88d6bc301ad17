"""
Raster -> cell-vector layout of the routing inputs (test infrastructure only).

Restates `DataLoader.vectorize` (xanthos/data_reader/data_load.py:416-425) and `sub2ind`
(xanthos/utils/math.py:40-50) as the loader uses them for the flow-distance, flow-direction and velocity
rasters (data_load.py:201-208, 392-413): the 280 x 720 DRT raster is flipped north-south, placed `skip` = 68
rows from the southern edge of a 360 x 720 map filled with -9999, flattened in Fortran order and sampled
at the cells' linear indices; values below `rep_val` are raised to it.
"""

import numpy as np


def sub2ind(shape, rows, cols):
    """Column-major linear index of (row, col) (math.py:40-50: ravel_multi_index(..., order='F'))."""
    return np.asarray(rows, dtype=np.int64) + np.asarray(cols, dtype=np.int64) * int(shape[0])


def vectorize(data, ngridrow, ngridcol, map_index, skip):
    """data_load.py:416-425."""
    new = np.zeros((ngridrow, ngridcol), dtype=float) - 9999
    for i in range(data.shape[0]):
        new[i + skip, :] = data[data.shape[0] - 1 - i, :]
    return new.reshape((ngridrow * ngridcol,), order='F')[map_index]


def load_routing_vector(raster, coords, ngridrow, ngridcol, skip=68, rep_val=None):
    """load_routing_data (data_load.py:392-413) for an in-memory raster; coords columns 4 / 3 = 1-based row / column
    of every cell (data_load.py:201-203)."""
    idx = sub2ind([ngridrow, ngridcol], coords[:, 4].astype(int) - 1, coords[:, 3].astype(int) - 1)
    v = vectorize(np.asarray(raster, dtype=float), ngridrow, ngridcol, idx, skip)
    if rep_val is not None:
        v[np.where(v < rep_val)[0]] = rep_val
    return v
