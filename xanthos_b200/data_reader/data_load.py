"""
Input staging - same attribute bag as xanthos/data_reader/data_load.py (DataLoader).

Hot-path inputs only: reference grids (:47-72), PET forcing and tables (:77-143), ABCD forcing
(:186-197), routing vectors (:200-211), calibration observations (:222-224).  File formats: .npy,
.csv, .txt (load_data, :343-390); MATLAB / NetCDF files are not read here (device-resident I/O of
the other formats is the "next" row f2 of SURVEY.md section 8).  Every forcing may also be passed
as an in-memory ndarray through `Xanthos.execute(args)` (load_to_array, :305-307).
"""

import logging
import os

import numpy as np

from .. import _cuda as C
from .._cuda import ValidationException  # noqa: F401  (same role as data_load.py:24)


def load_npy_pinned(fn):
    """
    np.load for the big [ncell, nmonths] float64 forcing files, read straight into a pinned host buffer of the
    recycling pool: the later upload then runs at PCIe speed instead of through the
    driver's pageable staging copy.  Anything else (small arrays, other dtypes, Fortran order, no CUDA device)
    goes through numpy.load.
    """
    try:
        with open(fn, 'rb') as f:
            major, _ = np.lib.format.read_magic(f)
            shape, fortran, dtype = (np.lib.format.read_array_header_1_0(f) if major == 1
                                     else np.lib.format.read_array_header_2_0(f))
            big = len(shape) == 2 and dtype == np.dtype('<f8') and not fortran and int(np.prod(shape)) >= (1 << 20)
            if big and C.device_available():
                arr = C.host_pool.as_array(C.host_pool.acquire(shape))
                view = memoryview(arr.reshape(-1)).cast('B')
                got = 0
                while got < len(view):
                    k = f.readinto(view[got:])
                    if not k:
                        raise IOError("Error: File {} is truncated".format(fn))
                    got += k
                return arr
    except (ValueError, OSError, AttributeError):
        pass
    return np.load(fn)


class DataLoader:
    """Load system-wide input data."""

    def __init__(self, config_obj):
        self.s = config_obj
        s = self.s

        # ---- reference grids (data_load.py:47-72) ---------------------------------------------------
        self.area = self.load_data(s.Area) * 0.01                      # ha -> km2
        self.coords = self.load_data(s.Coord)
        self.basin_ids = self.load_data(s.BasinIDs, 1).astype(int)
        self.basin_names = self._optional(s, 'BasinNames', lambda f: self.load_data(f))
        self.region_ids = self._optional(s, 'GCAMRegionIDs', lambda f: self.load_data(f, 1).astype(int))
        self.country_ids = self._optional(s, 'CountryIDs', lambda f: self.load_data(f, 1).astype(int))
        self.region_names = self._optional(s, 'GCAMRegionNames', self.get_region_names)     # data_load.py:63
        self.country_names = self._optional(s, 'CountryNames', self.get_country_names)      # data_load.py:69
        self.latitude = np.copy(self.coords[:, 2])
        self.lat_radians = np.radians(self.latitude)

        # ---- PET (data_load.py:77-143) -----------------------------------------------------------------
        if s.pet_module == 'hs':
            self.hs_tas = self.load_to_array(s.hs_tas)
            self.hs_tmin = self.load_to_array(s.hs_tmin)
            self.hs_tmax = self.load_to_array(s.hs_tmax)

        elif s.pet_module == 'pm':
            et = np.genfromtxt(s.pm_params, delimiter=',') if isinstance(s.pm_params, str) else np.asarray(s.pm_params)
            (self.cL, self.beta, self.rslimit, self.ae, self.be, self.Tminopen, self.Tminclose, self.VPDclose,
             self.VPDopen, self.RBLmin, self.RBLmax, self.rc, self.emiss) = [et[:, k] for k in range(13)]
            self.alpha = self._table(s.pm_alpha)
            self.lai = self._table(s.pm_lai)
            self.laimax = self._table(s.pm_laimax)
            self.laimin = self._table(s.pm_laimin)
            self.tair_load = self.load_to_array(s.pm_tas, 'pm_tas', nan_to_num=True)
            self.TMIN_load = self.load_to_array(s.pm_tmin, 'pm_tmin', nan_to_num=True)
            self.rhs_load = self.load_to_array(s.pm_rhs, 'pm_rhs', nan_to_num=True)
            self.wind_load = self.load_to_array(s.pm_wind, 'pm_wind', nan_to_num=True)
            self.rsds_load = self.load_to_array(s.pm_rsds, 'pm_rsds', nan_to_num=True)
            self.rlds_load = self.load_to_array(s.pm_rlds, 'pm_rlds', nan_to_num=True)
            # tairprev_load (data_load.py:127-129) is the temperature of the previous CELL; the CUDA
            # kernel reads tair[c - 1] directly, so no shifted copy is materialised here.
            lct = np.load(s.pm_lct) if isinstance(s.pm_lct, str) else np.asarray(s.pm_lct)
            self.lct_load = np.nan_to_num(lct)
            elev = np.load(s.pm_elev) if isinstance(s.pm_elev, str) else np.asarray(s.pm_elev)
            self.elev = np.nan_to_num(elev)

        elif s.pet_module == 'thornthwaite':
            self.tair = self.load_to_array(s.trn_tas, 'trn_tas', nan_to_num=True, warn_nan=True)

        elif s.pet_module == 'hargreaves':                         # data_load.py:77-84
            self.temp = self.load_to_array(s.TemperatureFile, var_name=s.TempVarName)
            self.dtr = self.load_to_array(s.DailyTemperatureRangeFile, var_name=s.DTRVarName, neg_to_zero=True)

        elif s.pet_module == 'none' and getattr(s, 'pet_file', None) is not None:
            self.pet_out = self.load_to_array(s.pet_file)

        # ---- runoff (data_load.py:186-197) ---------------------------------------------------------------
        if s.runoff_module == 'abcd':
            self.precip = self.load_to_array(s.PrecipitationFile, var_name=s.PrecipVarName, warn_nan=True)
            if s.TempMinFile is None:
                logging.info('TempMinFile variable not found for the ABCD runoff module; '
                             'Snowmelt will not be accounted for.')
                self.tmin = None
            else:
                self.tmin = self.load_to_array(s.TempMinFile, var_name=s.TempMinVarName, nan_to_num=True,
                                               warn_nan=True)

        elif s.runoff_module == 'gwam':                            # data_load.py:147-182
            self.precip = self.load_to_array(s.PrecipitationFile, var_name=s.PrecipVarName)
            self.max_soil_moist = self._vector(s.max_soil_moisture, 1)
            self.lakes_msm = self._pairs(s.lakes_msm)
            self.addit_water_msm = self._pairs(s.addit_water_msm)
            sm = np.array(self.max_soil_moist, dtype=float)
            sm[self.lakes_msm[:, 0]] = self.lakes_msm[:, 1]          # water bodies: 999 (load_soil_moisture :241-252)
            sm[self.addit_water_msm[:, 0]] = self.addit_water_msm[:, 1]
            self.grid_area = np.copy(self.area)
            self.soil_moisture = sm
            if str(s.HistFlag).lower() == "true":
                self.sm_prev = 0.5 * self.soil_moisture
            else:
                self.sm_prev = self.load_data(s.SavFile, 0, s.SavVarName)[:, -1]

        # ---- routing (data_load.py:200-211) ---------------------------------------------------------------
        if s.routing_module == 'mrtm':
            self.flow_dist = self.load_routing_data(s.flow_distance, rep_val=1000)
            self.flow_dir = self.load_routing_data(s.flow_direction)
            self.str_velocity = self.load_routing_data(s.strm_veloc, rep_val=0)
            self.instream_flow = np.zeros((s.ncell,), dtype=float)
            self.chs_prev = self.load_chs_data()

        if s.calibrate:
            self.cal_obs = self.load_data(s.cal_observed, 0)[:, [0, 3]]

    # ---------------------------------------------------------------------------------------------
    def _vector(self, f, header_num=0):
        return np.asarray(f, dtype=float).reshape(-1) if isinstance(f, np.ndarray) else self.load_data(f, header_num)

    def _pairs(self, f):
        """[k, 2] (1-based cell number, value) tables of water bodies -> 0-based int rows (data_load.py:154-160)."""
        a = np.asarray(f) if isinstance(f, np.ndarray) else self.load_data(f)
        a = np.atleast_2d(a).astype(int).reshape(-1, 2)
        a[:, 0] -= 1
        return a

    @staticmethod
    def _optional(s, attr, fn):
        f = getattr(s, attr, None)
        if f is None or (isinstance(f, str) and not os.path.isfile(f)):
            return None
        return fn(f)

    @staticmethod
    def _table(f):
        return np.genfromtxt(f, delimiter=',') if isinstance(f, str) else np.asarray(f)

    def load_chs_data(self):
        """Initial channel storage: zeros in historic mode, else last column of a previous run (:427-438)."""
        try:
            if str(self.s.HistFlag) == "True":
                return np.zeros((self.s.ncell,), dtype=float)
            return self.load_data(self.s.ChStorageFile, 0, self.s.ChStorageVarName)[:, -1]
        except AttributeError:
            return np.zeros((self.s.ncell,), dtype=float)

    @staticmethod
    def get_country_names(fn):
        """Second column of the headerless `index,name` file (data_load.py:275-279)."""
        with open(fn, 'r') as f:
            return np.array([ln.split(',')[1] for ln in f.read().splitlines() if ln.strip()])

    @staticmethod
    def get_region_names(fn):
        """First column of the region name file below its header line (data_load.py:281-286)."""
        with open(fn, 'r') as f:
            f.readline()
            return np.array([ln.split(',')[0] for ln in f.read().split('\n') if ln.strip()])

    def load_to_array(self, f, var_name=None, neg_to_zero=False, nan_to_num=False, warn_nan=False):
        """Load and validate a [ncell, nmonths] input (data_load.py:288-340)."""
        if isinstance(f, np.ndarray):
            arr = f
            f = 'in memory'
        else:
            arr = self.load_data(f, 0, var_name)
        if var_name is None:
            var_name = os.path.splitext(os.path.basename(f))[0]
        if neg_to_zero:
            arr[np.where(arr < 0)] = 0
        if warn_nan and np.any(np.isnan(arr)):
            logging.warning("NaNs found in input file {}".format(var_name))
        if nan_to_num:
            arr = np.nan_to_num(arr)
        return self.validate(arr, text=var_name)

    def validate(self, arr, text):
        """data_load.py:325-340."""
        err = "Error: Inconsistent {0} data grid size. Expecting size: {1}. Received size: {2}"
        if not arr.shape[0] == self.s.ncell:
            raise ValidationException(err.format(text, self.s.ncell, arr.shape[0]))
        if not arr.shape[1] == self.s.nmonths:
            raise ValidationException(err.format(text, self.s.nmonths, arr.shape[1]))
        return arr

    @staticmethod
    def load_data(fn, header_num=0, key=None):
        """data_load.py:343-390 (.npy / .txt / .csv)."""
        if isinstance(fn, np.ndarray):
            return fn
        if not os.path.isfile(fn):
            raise IOError("Error: File does not exist:", fn)
        if fn.endswith('.npy'):
            return load_npy_pinned(fn)
        if fn.endswith('.txt'):
            try:
                return np.genfromtxt(fn, delimiter=" ", skip_header=header_num, filling_values="0")
            except Exception:
                with open(fn, 'r') as f:
                    return np.array(f.read().splitlines())
        if fn.endswith('.csv'):
            return np.genfromtxt(fn, delimiter=",", skip_header=header_num, filling_values="0")
        if fn.endswith('.mat'):
            import scipy.io as sio
            return sio.loadmat(fn)[key]
        if fn.endswith('.nc'):
            # scipy.io.netcdf (the module the reference imports, data_load.py:376) is gone; the class remains
            import scipy.io as sio
            datagrp = sio.netcdf_file(fn, 'r', mmap=False)
            data = datagrp.variables[key][:].copy()
            datagrp.close()
            if data.dtype.byteorder == ">":                       # little-endian only, like the reference (:383-385)
                data = data.byteswap().view(data.dtype.newbyteorder('<'))
            return data
        raise RuntimeError("File {} has unrecognized extension".format(fn))

    def load_routing_data(self, fn, rep_val=None):
        """
        Routing vectors (data_load.py:392-425).  A [ncell] vector (in memory or .npy) is used as is;
        a 2-D raster is flipped, placed `skip` rows from the south edge and sampled in Fortran order
        like `vectorize` (:416-425).
        """
        fd = self.load_data(fn) if not isinstance(fn, np.ndarray) else fn
        fd = np.asarray(fd, dtype=float)
        if fd.ndim == 2 and fd.shape[0] != self.s.ncell:
            ilat = self.coords[:, 4].astype(int) - 1
            ilon = self.coords[:, 3].astype(int) - 1
            # the DRT rasters of the reference are 280 x 720 and start 68 rows from the south edge (:392)
            skip = 68 if (fd.shape[0] == 280 and self.s.ngridrow == 360) else (self.s.ngridrow - fd.shape[0]) // 2
            new = np.zeros((self.s.ngridrow, self.s.ngridcol), dtype=float) - 9999
            new[skip:skip + fd.shape[0], :] = fd[::-1, :]
            v = new[ilat, ilon]
        else:
            v = fd.reshape(-1).copy()
        if rep_val is not None:
            v[v < rep_val] = rep_val
        return v
