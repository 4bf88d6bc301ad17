"""
Output staging - counterpart of xanthos/data_writer/out_writer.py (OutWriter).

Kept: variable selection by `output_vars`, yearly aggregation (sum; mean for avgchflow, :103-112),
mm -> km3 conversion (not for avgchflow, :114-115), file naming `<var>_<unit>_<ProjectName>.<ext>`
(:119), 1-based cell index and YYYYMM / YYYY column names (:66-70, :123), basin / country / region
aggregates (:250-265).  Formats (out_writer.py:159-235): NetCDF classic (0, `scipy.io.netcdf_file` - the
`scipy.io.netcdf` module the reference imports no longer exists), csv (1), MATLAB (2), parquet (3, written with
pyarrow when fastparquet is absent; one row group, gzip) and npy (4).

When the array handed in is the one a CUDA stage returned, its device copy is still resident and
the yearly aggregation (xan_agg_to_year) and the basin / country / region sums (xan_group_sum) run on the
device.
"""

import logging
import os

import numpy as np

from .. import _cuda as C

FORMAT_NETCDF, FORMAT_CSV, FORMAT_MAT, FORMAT_PARQUET, FORMAT_NPY = 0, 1, 2, 3, 4
UNIT_MM_MTH, UNIT_KM3_MTH = 0, 1
NMONTHS = 12


def agg_to_year(arr, func='sum'):
    """[ncell, nmonths] -> [ncell, nyears] (out_writer.py:237-248)."""
    f = C.resident(arr)
    if f is not None:
        torch = C.torch_cuda()
        out = C.Field.empty(f.ncell, f.nmonths // NMONTHS, f.ld)
        C.check(C.lib().xan_agg_to_year(C.ptr(f.t), C.ptr(out.t), f.ncell, f.nmonths, f.ld, int(func == 'mean'),
                                        C.stream_ptr()))
        return C.remember(out.to_host(), out)
    # arrays that did not come from a CUDA stage (host-side writer only); pandas' groupby skips NaN
    a = np.asarray(arr)[:, :arr.shape[1] // NMONTHS * NMONTHS].reshape(arr.shape[0], -1, NMONTHS)
    ok = ~np.isnan(a)
    tot = np.where(ok, a, 0.0).sum(axis=2)
    if func == 'mean':
        with np.errstate(invalid='ignore', divide='ignore'):
            return tot / ok.sum(axis=2)
    return tot


class OutWriter:
    """Write out main Xanthos output variables."""

    def __init__(self, settings, grid_areas, all_outputs):
        self.output_names = [o for o in settings.output_vars if o in all_outputs.keys()]
        self.outputs = [all_outputs[o] for o in self.output_names]
        self.conversion_mm_km3 = np.asarray(grid_areas) / 1e6
        self.proj_name = settings.ProjectName
        self.out_folder = settings.OutputFolder
        self.out_format = settings.OutputFormat
        self.out_unit = settings.OutputUnit
        self.out_unit_str = settings.OutputUnitStr
        self.output_in_year = settings.OutputInYear
        years = range(settings.StartYear, settings.EndYear + 1)
        if self.output_in_year:
            self.time_steps = [str(y) for y in years]
        else:
            self.time_steps = ['{}{:02}'.format(y, m) for y in years for m in range(1, NMONTHS + 1)]
        if self.out_format not in (FORMAT_NETCDF, FORMAT_CSV, FORMAT_MAT, FORMAT_PARQUET, FORMAT_NPY):
            logging.warning("Unknown output format {}; writing output as .csv".format(self.out_format))
            self.out_format = FORMAT_CSV

    def get(self, varstr):
        return self.outputs[self.output_names.index(varstr)]

    def write(self):
        if not self.output_names:
            logging.debug("No valid output variables specified")
            return
        for i, var in enumerate(self.output_names):
            func, unit = ('mean', 'm3persec') if var == 'avgchflow' else ('sum', self.out_unit_str)
            if self.output_in_year:
                self.outputs[i] = agg_to_year(self.outputs[i], func)
            if self.out_unit == UNIT_KM3_MTH and var != 'avgchflow':
                self.outputs[i] = self.outputs[i] * self.conversion_mm_km3[:, None]
            filename = os.path.join(self.out_folder, '{}_{}_{}'.format(var, unit, self.proj_name))
            self.write_data(filename, var, self.outputs[i], self.time_steps)

    def write_data(self, filename, var, data, col_names=None, index_base=1):
        """index_base: first value of the id column (grid-cell outputs are 1-based, out_writer.py:123; the drought
        module hands over a fresh DataFrame whose index starts at 0, drought_stats.py:60-63)."""
        if self.out_format == FORMAT_NPY:
            np.save(filename + '.npy', data)
        elif self.out_format == FORMAT_MAT:
            import scipy.io as spio
            spio.savemat(filename + '.mat', {var: np.asarray(data)})                  # out_writer.py:181-183
        elif self.out_format == FORMAT_NETCDF:
            self.save_netcdf(filename, np.asarray(data), var)
        elif self.out_format == FORMAT_PARQUET:
            self.save_parquet(filename, np.asarray(data), col_names, index_base)
        else:
            if col_names is None:
                col_names = [str(k) for k in range(data.shape[1])]
            header = 'id,' + ','.join(col_names)
            ids = np.arange(index_base, data.shape[0] + index_base)[:, None]
            np.savetxt(filename + '.csv', np.hstack([ids, data]), delimiter=',', header=header, comments='',
                       fmt=['%d'] + ['%.17g'] * data.shape[1])

    def save_netcdf(self, filename, data, varstr):
        """NetCDF classic, float32 variable 'data' over ('index', 'month' | 'year') (out_writer.py:198-223)."""
        import scipy.io as spio
        datagrp = spio.netcdf_file(filename + '.nc', 'w')
        nrows, ncols = data.shape
        tdim = 'year' if self.output_in_year else 'month'
        datagrp.createDimension('index', nrows)
        datagrp.createDimension(tdim, ncols)
        griddata = datagrp.createVariable('data', 'f4', ('index', tdim))
        griddata.units = self.out_unit_str
        griddata.description = varstr + "_" + self.out_unit_str
        griddata[:, :] = data[:, :].copy()
        datagrp.close()

    def save_parquet(self, filename, data, col_names=None, index_base=1):
        """One gzip-compressed row group with the time steps as columns (out_writer.py:225-235; pyarrow instead of
        fastparquet's hive layout)."""
        import pyarrow as pa
        import pyarrow.parquet as pq
        if col_names is None:
            col_names = [str(k) for k in range(data.shape[1])]
        cols = {'id': np.arange(index_base, data.shape[0] + index_base)}
        cols.update({str(c): np.ascontiguousarray(data[:, k]) for k, c in enumerate(col_names)})
        pq.write_table(pa.table(cols), filename + '.parquet', compression='gzip', row_group_size=max(1, data.shape[0]))

    def write_aggregates(self, ref, data, basin, country, region):
        """Sum over basins / countries / regions (out_writer.py:250-265)."""
        for flag, ids, name in ((basin, getattr(ref, 'basin_ids', None), 'Basin_runoff'),
                                (country, getattr(ref, 'country_ids', None), 'Country_runoff'),
                                (region, getattr(ref, 'region_ids', None), 'GCAMRegion_runoff')):
            if not flag or ids is None:
                continue
            ids = np.asarray(ids).astype(int)
            nid = int(ids.max())
            f = C.resident(data)
            if f is not None:
                # the array a CUDA stage returned: summed where it lies in HBM (xan_group_sum, NaN skipped)
                from ..diagnostics.time_series import group_sum_device
                out = group_sum_device(ids, f.t).cpu().numpy()
            else:
                out = np.zeros((nid, data.shape[1]))
                for k in range(data.shape[1]):
                    out[:, k] = np.bincount(ids, weights=np.where(np.isnan(data[:, k]), 0.0, data[:, k]),
                                            minlength=nid + 1)[1:]
            fn = os.path.join(self.out_folder, '{}_{}_{}'.format(name, self.out_unit_str, self.proj_name))
            header = 'id,' + ','.join(self.time_steps)
            np.savetxt(fn + '.csv', np.hstack([np.arange(1, nid + 1)[:, None], out]), delimiter=',', header=header,
                       comments='', fmt=['%d'] + ['%.17g'] * out.shape[1])
