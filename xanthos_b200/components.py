"""
Components for use in model configurations - drop-in for xanthos/components.py.

Same class, attributes and step sequence as the reference (components.py:29-497): `Components`
owns a `DataLoader` and the six output arrays PET, AET, Q, Sav, ChStorage, Avg_ChFlow
[ncell, nmonths]; `simulation()` runs PET -> runoff -> routing by dispatching to the plug-in modules
bound in `import_core()`.  The plug-ins are the CUDA-backed modules of this package; each returns
host arrays (as the reference's do) and leaves a device copy resident, so the next stage starts
from HBM.  Routing runs the reference's two month loops (components.py:273-294) as ONE kernel
launch (`routing_mod.route`).

Post-processing: drought statistics, accessible water and the aggregated time series run on the
resident fields; hydropower and the diagnostics against external data
sets are outside the scope and are logged no-ops.
"""

import logging
import time

import numpy as np

from . import _cuda as C
from .utils import general as helper
from .calibrate import calibrate_abcd as calib_mod
from .data_writer.out_writer import OutWriter
from .data_reader.data_load import DataLoader
from .drought.drought_stats import DroughtStats
from .accessible.accessible import AccessibleWater
from .diagnostics.time_series import TimeSeriesPlot

pet_mod = None
runoff_mod = None
routing_mod = None


class Components:
    """Components for use in model configurations."""

    def __init__(self, config):
        self.s = config
        self.import_core()
        self.data = DataLoader(config)
        self.yr_imth_dys = helper.set_month_arrays(self.s.nmonths, self.s.StartYear, self.s.EndYear)

        if self.s.pet_module == 'hargreaves':                      # components.py:57-63
            self.pet_out = None
            self.solar_dec, self.dr = helper.calc_sinusoidal_factor(self.yr_imth_dys)
        if self.s.runoff_module == 'gwam':                         # components.py:75-77
            self.soil_moisture = self.data.soil_moisture
            self.sm_prev = self.data.sm_prev
        self.P = self.T = self.D = None
        self.mth_solar_dec = self.mth_dr = self.mth_days = None
        self.pet_t = None

        if self.s.routing_module == 'mrtm':
            self.flow_dist = None
            self.flow_dir = None
            self.instream_flow = None
            self.str_velocity = None
            self.dsid = None
            self.upid = None
            self.um = None
            self.routing_timestep_hours = 3 * 3600      # components.py:91 (seconds, despite the name)
            self.chs_prev = None

        shp = (self.s.ncell, self.s.nmonths)
        self.PET = np.zeros(shape=shp)
        self.AET = np.zeros(shape=shp)
        self.Q = np.zeros(shape=shp)
        self.Sav = np.zeros(shape=shp)
        self.ChStorage = np.zeros(shape=shp)
        self.Avg_ChFlow = np.zeros(shape=shp)
        self.q = None
        self.ac = None

    def import_core(self):
        """Bind the PET / runoff / routing plug-in modules (components.py:114-142)."""
        global pet_mod, runoff_mod, routing_mod
        if self.s.pet_module == 'hargreaves':
            from .pet import hargreaves as pet_mod
        elif self.s.pet_module == 'hs':
            from .pet import hargreaves_samani as pet_mod
        elif self.s.pet_module == 'pm':
            from .pet import penman_monteith as pet_mod
        elif self.s.pet_module == 'thornthwaite':
            from .pet import thornthwaite as pet_mod
        if self.s.runoff_module == 'gwam':
            from .runoff import gwam as runoff_mod
        elif self.s.runoff_module == 'abcd':
            from .runoff import abcd as runoff_mod
        if self.s.routing_module == 'mrtm':
            from .routing import mrtm as routing_mod

    def prep_arrays(self, nm=None):
        """Climate arrays of one month, or of all months (components.py:143-157)."""
        if nm is None:
            self.P = np.copy(self.data.precip)                      # keep nan in P
            self.T = np.nan_to_num(self.data.temp)
            self.D = np.nan_to_num(self.data.dtr)
        else:
            self.P = np.copy(self.data.precip[:, nm])
            self.T = np.nan_to_num(self.data.temp[:, nm])
            self.D = np.nan_to_num(self.data.dtr[:, nm])

    def prep_pet(self, nm):
        """Per-month scalars and arrays of the Hargreaves PET (components.py:159-186)."""
        if self.s.pet_module == 'hargreaves':
            self.mth_solar_dec = np.copy(self.solar_dec[nm])
            self.mth_dr = np.copy(self.dr[nm])
            self.mth_days = np.copy(self.yr_imth_dys[nm, 2])
            self.mth_temp_pet = np.nan_to_num(self.T)
            self.mth_dtr_pet = np.nan_to_num(self.D)

    def calculate_pet(self):
        """Monthly potential evapotranspiration (components.py:189-210)."""
        if self.s.pet_module == 'hargreaves':
            return pet_mod.calculate_pet(self.mth_temp_pet, self.mth_dtr_pet, self.data.lat_radians,
                                         self.mth_solar_dec, self.mth_dr, self.mth_days)
        elif self.s.pet_module == 'hs':
            return pet_mod.execute(self.s, self.data)
        elif self.s.pet_module == 'pm':
            return pet_mod.run_pmpet(self.data, self.s.ncell, self.s.pm_nlcs, self.s.StartYear, self.s.EndYear,
                                     self.s.pm_water_idx, self.s.pm_snow_idx, self.s.pm_lc_years)
        elif self.s.pet_module == 'thornthwaite':
            return pet_mod.execute(self.data.tair, self.data.lat_radians, self.s.StartYear, self.s.EndYear)
        elif self.s.pet_module == 'none':
            return self.data.pet_out

    def calculate_runoff(self, step_num=None, pet=None):
        """Runoff (components.py:212-247); ABCD iterates internally, GWAM is one month per call."""
        if self.s.runoff_module == 'gwam':
            rg = runoff_mod.runoffgen(self.pet_t, self.P, self.s, self.soil_moisture, self.sm_prev)
            self.PET[:, step_num], self.AET[:, step_num], self.Q[:, step_num], self.Sav[:, step_num] = rg
        elif self.s.runoff_module == 'abcd':
            rg = runoff_mod.abcd_execute(n_basins=self.s.n_basins, basin_ids=self.data.basin_ids,
                                         pet=pet, precip=self.data.precip, tmin=self.data.tmin,
                                         calib_file=self.s.calib_file, n_months=self.s.nmonths,
                                         spinup_steps=self.s.runoff_spinup, jobs=self.s.ro_jobs)
            self.PET, self.AET, self.Q, self.Sav = rg
        else:
            if getattr(self.s, 'alt_runoff', None) is not None:
                self.Q = np.load(self.s.alt_runoff)

    def calculate_routing(self, runoff):
        """
        Routing (components.py:249-296): topology, spin-up over the first `routing_spinup` months,
        then all months.  Returns Avg_ChFlow like the reference.
        """
        if self.s.routing_module == 'mrtm':
            self.flow_dist = self.data.flow_dist
            self.flow_dir = self.data.flow_dir
            self.instream_flow = self.data.instream_flow
            self.str_velocity = self.data.str_velocity
            self.chs_prev = self.data.chs_prev

            if self.um is None:
                self.dsid = routing_mod.downstream(self.data.coords, self.flow_dir, self.s)
                self.upid = routing_mod.upstream(self.data.coords, self.dsid, self.s)
                self.um = routing_mod.upstream_genmatrix(self.upid)

            sr = routing_mod.route(self.um, runoff, self.flow_dist, self.str_velocity, self.data.area,
                                   self.yr_imth_dys[:, 2], self.routing_timestep_hours, self.s.routing_spinup,
                                   chs_prev=self.chs_prev)
            self.ChStorage, self.Avg_ChFlow, self.instream_flow = sr
            C.host_sync()                                   # the copy of ChStorage may still be in flight
            self.chs_prev = np.copy(self.ChStorage[:, -1])
            return self.Avg_ChFlow

    def simulation(self, run_pet, run_runoff, run_routing, pet_num_steps=0, runoff_num_steps=0, routing_num_steps=0,
                   notify='simulation'):
        """Run PET -> runoff -> routing (components.py:298-385)."""
        if self.s.calibrate:
            self.calibrate()
            return

        logging.info("---{} in progress...".format(notify))
        t0 = time.time()
        with C.async_host():      # D2H of each stage's results overlaps the next stage; complete on exit
            if pet_num_steps > 0 or runoff_num_steps > 0:
                self._simulation_stepwise(run_pet, run_runoff, run_routing, pet_num_steps, runoff_num_steps, notify)
            else:
                self._simulation(run_pet, run_runoff, run_routing)
        logging.info("---{0} has finished successfully: {1} seconds ---".format(notify, time.time() - t0))

    def _simulation_stepwise(self, run_pet, run_runoff, run_routing, pet_num_steps, runoff_num_steps, notify):
        """
        The month loops of the reference (components.py:329-340 Hargreaves, :358-366 GWAM) as two launches:
        PET of every month is independent, and GWAM's soil moisture feedback (sm_prev) stays in registers.
        In the 'Spin Up' pass of ConfigRunner (configurations.py:106-113) the reference also fills the first
        months of PET/AET/Q/Sav and routes them; every one of those values is overwritten by the simulation
        pass, only `sm_prev` survives - so only `sm_prev` is produced here.
        """
        spin_pass = notify == 'Spin Up'
        pet_f = None
        if run_pet and self.s.pet_module == 'hargreaves':
            logging.info("\tProcessing PET...")
            t = time.time()
            if getattr(self, '_pet_series', None) is None:
                self._pet_series = pet_mod.series_device(self.data.temp, self.data.dtr, self.data.lat_radians,
                                                         self.solar_dec, self.dr, self.yr_imth_dys[:, 2])
            pet_f = self._pet_series
            logging.info("\tPET processed in {} seconds---".format(time.time() - t))
        elif run_pet or self.s.pet_module == 'none':
            pet_f = C.as_field(self.calculate_pet())

        if run_runoff and self.s.runoff_module == 'gwam':
            logging.info("\tProcessing Runoff...")
            t = time.time()
            res = runoff_mod.run_device(pet_f, self.data.precip, self.soil_moisture, self.sm_prev,
                                        n_months=runoff_num_steps, spinup_months=0,
                                        want=() if spin_pass else ('aet', 'q', 'sav'))
            self.sm_prev = res['sm_last'].cpu().numpy()             # components.py:366
            if not spin_pass:
                self.PET = C.remember(pet_f.to_host(), pet_f)
                self.AET = C.remember(res['aet'].to_host(), res['aet'])
                self.Q = C.remember(res['q'].to_host(), res['q'])
                self.Sav = C.remember(res['sav'].to_host(), res['sav'])
            logging.info("\tRunoff processed in {} seconds---".format(time.time() - t))
        elif run_runoff:
            self.calculate_runoff(pet=C.remember(pet_f.to_host(), pet_f))

        if run_routing and not spin_pass:
            logging.info("\tProcessing Routing...")
            t = time.time()
            self.calculate_routing(self.Q)
            logging.info("\tRouting processed in {} seconds---".format(time.time() - t))

    def _simulation(self, run_pet, run_runoff, run_routing):
        if run_pet:
            logging.info("\tProcessing PET...")
            t = time.time()
            pet_out = self.calculate_pet()
            logging.info("\tPET processed in {} seconds---".format(time.time() - t))
        else:
            pet_out = self.calculate_pet()

        if run_runoff and self.s.runoff_module == 'abcd':
            # Upload the runoff forcing on the side stream while PET computes.  Queued AFTER the uploads of the PET
            # module (the copy engine serves them in submission order): PET starts as soon as ITS forcing is there
            # and runs under the upload of precipitation and minimum temperature, instead of behind it.
            C.prefetch(self.data.precip)
            C.prefetch(self.data.tmin)

        if run_runoff:
            logging.info("\tProcessing Runoff...")
            t = time.time()
            self.calculate_runoff(pet=pet_out)
            logging.info("\tRunoff processed in {} seconds---".format(time.time() - t))

        if run_routing:
            logging.info("\tProcessing Routing...")
            t = time.time()
            self.calculate_routing(self.Q)
            logging.info("\tRouting processed in {} seconds---".format(time.time() - t))

    # ---- post-processing: outside the hot path -------------------
    def _skipped(self, flag, name):
        if flag:
            logging.warning("---%s is not part of xanthos_b200 (post-processing, out of scope); skipped", name)

    def drought(self):
        """Drought statistics on the resident fields (components.py:391-400)."""
        if self.s.CalculateDroughtStats:
            logging.info("---Start Drought Statistics:")
            t0 = time.time()
            DroughtStats(self.s, self.Q, self.Sav)
            logging.info("---Drought Statistics has finished successfully: %s seconds ------" % (time.time() - t0))

    def accessible_water(self):
        """Accessible water per basin (components.py:402-410)."""
        if self.s.CalculateAccessibleWater:
            logging.info("---Start Accessible Water:")
            t0 = time.time()
            AccessibleWater(self.s, self.data, self.Q)
            logging.info("---Accessible Water has finished successfully: %s seconds ------" % (time.time() - t0))

    def hydropower_potential(self):
        self._skipped(self.s.CalculateHydropowerPotential, 'Hydropower Potential')

    def hydropower_actual(self):
        self._skipped(self.s.CalculateHydropowerActual, 'Hydropower Actual')

    def diagnostics(self):
        self._skipped(self.s.PerformDiagnostics, 'Diagnostics')

    def plots(self):
        """Aggregated series of components.py:476-484 (written as .csv; no matplotlib here)."""
        if self.s.CreateTimeSeriesPlot:
            logging.info("---Creating Time Series Plots:")
            t0 = time.time()
            TimeSeriesPlot(self.s, self.q if self.q is not None else self.Q,
                           self.ac if self.ac is not None else self.Avg_ChFlow, self.data)
            logging.info("---Time Series have finished successfully: %s seconds ------" % (time.time() - t0))

    def output_simulation(self):
        """Write outputs (components.py:441-474)."""
        logging.info("---Output simulation results:")
        t0 = time.time()
        all_outputs = {'pet': self.PET, 'aet': self.AET, 'q': self.Q, 'soilmoisture': self.Sav,
                       'avgchflow': self.Avg_ChFlow}
        output_writer = OutWriter(self.s, self.data.area, all_outputs)
        output_writer.write()
        try:
            self.q = output_writer.get('q')
        except ValueError:
            self.q = self.Q
        try:
            self.ac = output_writer.get('avgchflow')
        except ValueError:
            self.ac = self.Avg_ChFlow
        output_writer.write_aggregates(self.data, self.q, self.s.AggregateRunoffBasin, self.s.AggregateRunoffCountry,
                                       self.s.AggregateRunoffGCAMRegion)
        logging.info("---Output finished: %s seconds ---" % (time.time() - t0))

    def calibrate(self):
        """Calibration of the ABCD parameters (components.py:486-497)."""
        logging.info("---Processing PET...")
        t = time.time()
        pet_out = self.calculate_pet()
        logging.info("---PET processed in {} seconds---".format(time.time() - t))
        logging.info("---Running calibration:")
        calib_mod.calibrate_all(settings=self.s, data=self.data, pet=pet_out, router_function=self.calculate_routing)
