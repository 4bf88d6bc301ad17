"""
Components for use in model configurations - drop-in for xanthos/components.py.

Same class, attributes and step sequence as the reference (components.py:29-497): `Components`
owns a `DataLoader` and the six output arrays PET, AET, Q, Sav, ChStorage, Avg_ChFlow
[ncell, nmonths]; `simulation()` runs PET -> runoff -> routing by dispatching to the plug-in modules
bound in `import_core()`.  The plug-ins are the CUDA-backed modules of this package; each returns
host arrays (as the reference's do) and leaves a device copy resident, so the next stage starts
from HBM.  Routing runs the reference's two month loops (components.py:273-294) as ONE kernel
launch (`routing_mod.route`).

Post-processing methods (drought, accessible water, hydropower, diagnostics, plots) are outside
the hot path and are logged no-ops here.
"""

import logging
import time

import numpy as np

from . import _cuda as C
from .utils import general as helper
from .calibrate import calibrate_abcd as calib_mod
from .data_writer.out_writer import OutWriter
from .data_reader.data_load import DataLoader

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
        if self.s.pet_module == 'hs':
            from .pet import hargreaves_samani as pet_mod
        elif self.s.pet_module == 'pm':
            from .pet import penman_monteith as pet_mod
        elif self.s.pet_module == 'thornthwaite':
            from .pet import thornthwaite as pet_mod
        if self.s.runoff_module == 'abcd':
            from .runoff import abcd as runoff_mod
        if self.s.routing_module == 'mrtm':
            from .routing import mrtm as routing_mod

    def calculate_pet(self):
        """Monthly potential evapotranspiration (components.py:189-210)."""
        if self.s.pet_module == 'hs':
            return pet_mod.execute(self.s, self.data)
        elif self.s.pet_module == 'pm':
            return pet_mod.run_pmpet(self.data, self.s.ncell, self.s.pm_nlcs, self.s.StartYear, self.s.EndYear,
                                     self.s.pm_water_idx, self.s.pm_snow_idx, self.s.pm_lc_years)
        elif self.s.pet_module == 'thornthwaite':
            return pet_mod.execute(self.data.tair, self.data.lat_radians, self.s.StartYear, self.s.EndYear)
        elif self.s.pet_module == 'none':
            return self.data.pet_out

    def calculate_runoff(self, step_num=None, pet=None):
        """Runoff (components.py:212-247); ABCD iterates internally."""
        if self.s.runoff_module == 'abcd':
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
            self._simulation(run_pet, run_runoff, run_routing)
        logging.info("---{0} has finished successfully: {1} seconds ---".format(notify, time.time() - t0))

    def _simulation(self, run_pet, run_runoff, run_routing):
        if run_runoff and self.s.runoff_module == 'abcd':
            # start uploading the runoff forcing while PET computes
            C.prefetch(self.data.precip)
            C.prefetch(self.data.tmin)

        if run_pet:
            logging.info("\tProcessing PET...")
            t = time.time()
            pet_out = self.calculate_pet()
            logging.info("\tPET processed in {} seconds---".format(time.time() - t))
        else:
            pet_out = self.calculate_pet()

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
        self._skipped(self.s.CalculateDroughtStats, 'Drought Statistics')

    def accessible_water(self):
        self._skipped(self.s.CalculateAccessibleWater, 'Accessible Water')

    def hydropower_potential(self):
        self._skipped(self.s.CalculateHydropowerPotential, 'Hydropower Potential')

    def hydropower_actual(self):
        self._skipped(self.s.CalculateHydropowerActual, 'Hydropower Actual')

    def diagnostics(self):
        self._skipped(self.s.PerformDiagnostics, 'Diagnostics')

    def plots(self):
        self._skipped(self.s.CreateTimeSeriesPlot, 'Time Series Plots')

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
