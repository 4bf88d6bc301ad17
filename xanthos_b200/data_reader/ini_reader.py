"""
Configuration reader - same attribute bag as xanthos/data_reader/ini_reader.py (ConfigReader).

The reference parses the .ini with `configobj`, which is not a dependency here: `parse_ini` below
reads the subset of the format the reference's files use (nested sections `[a]` / `[[b]]`,
`key = value`, `#` comments, comma separated lists, optional quotes).  Attribute names, defaults,
path joins and error types follow ini_reader.py:27-196 (project), :198-300 (PET), :302-390
(runoff), :392-425 (routing), :427-440 (reference files), :504-519 (calibration).

Only the hot-path sections are interpreted (PET / Runoff / Routing / Calibrate); the optional
post-processing sections are parsed into `self.sections` but not configured.

Extension: `[Project]` may carry `ncell`, `ngridrow`, `ngridcol` (defaults 67420 / 360 / 720, the
constants of ini_reader.py:117-119) so that reduced worlds can be run in tests.
"""

import logging
import os


class ValidationException(Exception):
    """Custom exception for invalid Xanthos inputs (ini_reader.py:17)."""


def _convert(value):
    v = value.strip()
    if ',' in v:
        return [p.strip().strip('"\'') for p in v.split(',') if p.strip() != '']
    return v.strip('"\'')


def parse_ini(path):
    """Nested dict of the sections of a configobj-style ini file."""
    root = {}
    stack = [root]
    with open(path, 'r') as f:
        for raw in f:
            line = raw.split('#', 1)[0].rstrip() if not raw.lstrip().startswith('#') else ''
            if not line.strip():
                continue
            s = line.strip()
            if s.startswith('['):
                depth = len(s) - len(s.lstrip('['))
                name = s.strip('[]').strip()
                if depth > len(stack):
                    raise ValidationException("Section '{}' nested too deeply in {}".format(name, path))
                del stack[depth:]
                new = {}
                stack[-1][name] = new
                stack.append(new)
            elif '=' in s:
                k, v = s.split('=', 1)
                stack[-1][k.strip()] = _convert(v)
    return root


class ConfigReader:
    """Read the Xanthos configuration .ini file (attribute names identical to the reference)."""

    def __init__(self, ini):
        c = parse_ini(ini) if not isinstance(ini, dict) else ini
        self.sections = c
        p = c['Project']

        self.root = p['RootDir']
        self.ProjectName = p['ProjectName']
        self.OutputNameStr = p['ProjectName']
        self.InputFolder = os.path.join(self.root, p['InputFolder'])
        self.OutDir = self.create_dir(os.path.join(self.root, p['OutputFolder']))
        self.OutputFolder = self.create_dir(os.path.join(self.OutDir, self.ProjectName))

        ref = 'RefDir' in p
        if ref:
            self.Reference = os.path.join(self.InputFolder, p['RefDir'])
        pet_config = c.get('PET', False)
        if pet_config and 'pet_dir' in p:
            self.PET = os.path.join(self.InputFolder, p['pet_dir'])
        elif pet_config:
            pet_config = False
        runoff_config = c.get('Runoff', False)
        if runoff_config and 'RunoffDir' in p:
            self.RunoffDir = os.path.join(self.InputFolder, p['RunoffDir'])
        elif runoff_config:
            runoff_config = False
        routing_config = c.get('Routing', False)
        if routing_config and 'RoutingDir' in p:
            self.RoutingDir = os.path.join(self.InputFolder, p['RoutingDir'])
        elif routing_config:
            routing_config = False
        calibration_config = c.get('Calibrate', False)
        drought_config = c.get('Drought', False)                       # ini_reader.py:85-88
        acc_water_config = c.get('AccessibleWater', False)              # ini_reader.py:90-94
        if acc_water_config and 'AccWatDir' in p:
            self.AccWatDir = os.path.join(self.InputFolder, p['AccWatDir'])
        elif acc_water_config:
            acc_water_config = False
        timeseries_config = c.get('TimeSeriesPlot', False)

        # project level settings (ini_reader.py:117-139)
        self.ncell = int(p.get('ncell', 67420))
        self.ngridrow = int(p.get('ngridrow', 360))
        self.ngridcol = int(p.get('ngridcol', 720))
        self.n_basins = int(p['n_basins'])
        self.HistFlag = p['HistFlag']
        self.StartYear = int(p['StartYear'])
        self.EndYear = int(p['EndYear'])
        self.output_vars = p['output_vars']
        self.OutputFormat = int(p['OutputFormat'])
        self.OutputUnit = int(p['OutputUnit'])
        self.OutputInYear = int(p['OutputInYear'])
        self.AggregateRunoffBasin = int(p['AggregateRunoffBasin'])
        self.AggregateRunoffCountry = int(p['AggregateRunoffCountry'])
        self.AggregateRunoffGCAMRegion = int(p['AggregateRunoffGCAMRegion'])
        self.PerformDiagnostics = int(p['PerformDiagnostics'])
        self.CreateTimeSeriesPlot = int(p['CreateTimeSeriesPlot'])
        self.CalculateDroughtStats = int(p['CalculateDroughtStats'])
        self.CalculateAccessibleWater = int(p['CalculateAccessibleWater'])
        self.CalculateHydropowerPotential = int(p['CalculateHydropowerPotential'])
        self.CalculateHydropowerActual = int(p['CalculateHydropowerActual'])
        self.calibrate = int(p['Calibrate'])

        self.nmonths = (self.EndYear - self.StartYear + 1) * 12
        self.OutputUnitStr = '{}per{}'.format(('mm', 'km3')[self.OutputUnit], ('month', 'year')[self.OutputInYear])
        self.output_vars = [self.output_vars] if not isinstance(self.output_vars, list) else self.output_vars

        self.configure_pet(pet_config)
        self.configure_runoff(runoff_config)
        self.configure_routing(routing_config)

        self.mod_cfg = '{0}_{1}_{2}'.format(self.pet_module, self.runoff_module, self.routing_module)
        if self.mod_cfg == 'none_none_none':
            raise ValidationException('No PET, Runoff, or Routing model selected.')

        self.configure_reference_data(ref)

        if drought_config and self.CalculateDroughtStats:
            self.configure_drought_stats(drought_config)
        if acc_water_config and self.CalculateAccessibleWater:
            self.configure_acc_water(acc_water_config)
        if timeseries_config and self.CreateTimeSeriesPlot:
            self.configure_timeseries_plot(timeseries_config)

        for flag, name in ((self.PerformDiagnostics, 'Diagnostics'),
                           (self.CalculateHydropowerPotential, 'HydropowerPotential'),
                           (self.CalculateHydropowerActual, 'HydropowerActual')):
            if flag:
                logging.warning("Post-processing module '%s' is outside the scope of xanthos_b200 and will be "
                                "skipped (see DESIGN.md).", name)

        if calibration_config and self.calibrate:
            self.configure_calibration(calibration_config)

    # ---------------------------------------------------------------------------------------------
    def configure_pet(self, pet_config):
        """ini_reader.py:198-300."""
        if not pet_config:
            self.pet_module = 'none'
            self.pet_file = None
            return

        self.pet_module = pet_config['pet_module'].lower()

        if self.pet_module == 'hs':
            m = pet_config['hargreaves-samani']
            self.pet_dir = os.path.join(self.PET, m['pet_dir'])
            self.hs_tas = os.path.join(self.pet_dir, m['hs_tas'])
            self.hs_tmin = os.path.join(self.pet_dir, m['hs_tmin'])
            self.hs_tmax = os.path.join(self.pet_dir, m['hs_tmax'])

        elif self.pet_module == 'pm':
            m = pet_config['penman-monteith']
            self.pet_dir = os.path.join(self.PET, m['pet_dir'])
            self.pm_tas = os.path.join(self.pet_dir, m['pm_tas'])
            self.pm_tmin = os.path.join(self.pet_dir, m['pm_tmin'])
            self.pm_rhs = os.path.join(self.pet_dir, m['pm_rhs'])
            self.pm_rlds = os.path.join(self.pet_dir, m['pm_rlds'])
            self.pm_rsds = os.path.join(self.pet_dir, m['pm_rsds'])
            self.pm_wind = os.path.join(self.pet_dir, m['pm_wind'])
            self.pm_lct = os.path.join(self.pet_dir, m['pm_lct'])
            self.pm_nlcs = int(m['pm_nlcs'])
            self.pm_water_idx = int(m['pm_water_idx'])
            self.pm_snow_idx = int(m['pm_snow_idx'])
            yrs = m['pm_lc_years']
            self.pm_lc_years = [int(i) for i in (yrs if isinstance(yrs, list) else [yrs])]
            self.pm_params = os.path.join(self.pet_dir, 'gcam_ET_para.csv')
            self.pm_alpha = os.path.join(self.pet_dir, 'gcam_albedo.csv')
            self.pm_lai = os.path.join(self.pet_dir, 'gcam_lai.csv')
            self.pm_laimin = os.path.join(self.pet_dir, 'gcam_laimin.csv')
            self.pm_laimax = os.path.join(self.pet_dir, 'gcam_laimax.csv')
            self.pm_elev = os.path.join(self.pet_dir, 'elev.npy')

        elif self.pet_module == 'thornthwaite':
            m = pet_config['thornthwaite']
            self.pet_dir = os.path.join(self.PET, m['pet_dir'])
            self.trn_tas = os.path.join(self.pet_dir, m['trn_tas'])

        elif self.pet_module == 'none':
            try:
                self.pet_file = pet_config['pet_file']
            except KeyError:
                raise ValidationException(
                    "USAGE: Must provide a pet_file variable in the PET config section that "
                    "contains the full path to an input PET file if not using an existing module.")

        elif self.pet_module == 'hargreaves':                      # ini_reader.py:216-243
            m = pet_config['hargreaves']
            self.pet_dir = os.path.join(self.PET, m['pet_dir'])
            try:
                self.TemperatureFile = os.path.join(self.pet_dir, m['TemperatureFile'])
            except KeyError:
                logging.exception("File path not provided for the TemperatureFile "
                                  "variable in the PET section of the config file.")
                raise
            self.TempVarName = m.get('TempVarName')
            try:
                self.DailyTemperatureRangeFile = os.path.join(self.pet_dir, m['DailyTemperatureRangeFile'])
            except KeyError:
                logging.exception("File path not provided for the DailyTemperatureRangeFile "
                                  "variable in the PET section of the config file.")
                raise
            self.DTRVarName = m.get('DTRVarName')
        else:
            raise ValidationException("ERROR: PET module '{0}' not found. Please check "
                                      "spelling and try again.".format(self.pet_module))

    def configure_runoff(self, runoff_config):
        """ini_reader.py:302-390."""
        if not runoff_config:
            self.runoff_module = 'none'
            return
        self.runoff_module = runoff_config['runoff_module'].lower()

        if self.runoff_module == 'abcd':
            m = runoff_config['abcd']
            self.ro_model_dir = os.path.join(self.RunoffDir, m['runoff_dir'])
            self.calib_file = os.path.join(self.ro_model_dir, m['calib_file'])
            self.runoff_spinup = int(m['runoff_spinup'])
            self.ro_jobs = int(m['jobs'])
            try:
                self.PrecipitationFile = m['PrecipitationFile']
            except KeyError:
                logging.exception("File path not provided for the PrecipitationFile "
                                  "variable in the ABCD runoff section of the config file.")
                raise
            self.PrecipVarName = m.get('PrecipVarName')
            self.TempMinFile = m.get('TempMinFile')
            self.TempMinVarName = m.get('TempMinVarName')

        elif self.runoff_module == 'none':
            pass
        elif self.runoff_module == 'gwam':                         # ini_reader.py:311-345
            m = runoff_config['gwam']
            self.ro_model_dir = os.path.join(self.RunoffDir, m['runoff_dir'])
            self.runoff_spinup = int(m['runoff_spinup'])
            self.max_soil_moisture = os.path.join(self.ro_model_dir, m['max_soil_moisture'])
            self.lakes_msm = os.path.join(self.ro_model_dir, m['lakes_msm'])
            self.addit_water_msm = os.path.join(self.ro_model_dir, m['addit_water_msm'])
            self.ChStorageFile = None
            self.ChStorageVarName = None
            self.SavFile = None
            self.SavVarName = None
            if str(self.HistFlag) == 'False':
                try:
                    self.ChStorageFile = m['ChStorageFile']
                    self.ChStorageVarName = m['ChStorageVarName']
                    self.SavFile = m['SavFile']
                    self.SavVarName = m['SavVarName']
                except KeyError:
                    raise ValidationException("Error: ChStorageFile and ChStorageVarName "
                                              "are not defined for Future Mode.")
            try:
                self.PrecipitationFile = os.path.join(self.ro_model_dir, m['PrecipitationFile'])
            except KeyError:
                logging.exception("File path not provided for the PrecipitationFile variable "
                                  "in the GWAM runoff section of the config file.")
                raise
            self.PrecipVarName = m.get('PrecipVarName')
        else:
            raise ValidationException("ERROR: Runoff module '{0}' not found. Please check "
                                      "spelling and try again.".format(self.runoff_module))

    def configure_routing(self, routing_config):
        """ini_reader.py:392-425."""
        if not routing_config:
            self.routing_module = 'none'
            return
        self.routing_module = routing_config['routing_module'].lower()

        if self.routing_module == 'mrtm':
            m = routing_config[self.routing_module]
            self.rt_model_dir = os.path.join(self.RoutingDir, m['routing_dir'])
            self.strm_veloc = os.path.join(self.rt_model_dir, m['channel_velocity'])
            self.flow_distance = os.path.join(self.rt_model_dir, m['flow_distance'])
            self.flow_direction = os.path.join(self.rt_model_dir, m['flow_direction'])
            try:
                self.routing_spinup = int(m['routing_spinup'])
            except KeyError:
                self.routing_spinup = self.nmonths
            try:
                self.alt_runoff = self.custom_runoff(m['alt_runoff'])
            except KeyError:
                self.alt_runoff = None
        elif self.routing_module == 'none':
            pass
        else:
            raise ValidationException("ERROR: Routing module '{0}' not found. Please check "
                                      "spelling and try again.".format(self.routing_module))

    def configure_reference_data(self, ref):
        """ini_reader.py:427-440."""
        if ref:
            self.Area = os.path.join(self.Reference, 'Grid_Areas_ID.csv')
            self.Coord = os.path.join(self.Reference, 'coordinates.csv')
            self.BasinIDs = os.path.join(self.Reference, 'basin.csv')
            self.BasinNames = os.path.join(self.Reference, 'BasinNames235.txt')
            self.GCAMRegionIDs = os.path.join(self.Reference, 'region32_grids.csv')
            self.GCAMRegionNames = os.path.join(self.Reference, 'Rgn32Names.csv')
            self.CountryIDs = os.path.join(self.Reference, 'country.csv')
            self.CountryNames = os.path.join(self.Reference, 'country-names.csv')
        else:
            logging.warning('No reference data selected for use.')

    def configure_drought_stats(self, drought_config):
        """ini_reader.py:460-471."""
        self.drought_var = drought_config['drought_var']
        self.drought_thresholds = drought_config.get('drought_thresholds')  # optional
        if self.drought_thresholds is None:
            self.threshold_nper = int(drought_config['threshold_nper'])
            self.threshold_start_year = int(drought_config['threshold_start_year'])
            self.threshold_end_year = int(drought_config['threshold_end_year'])
            if (self.StartYear > self.threshold_start_year) or (self.EndYear < self.threshold_end_year):
                raise ValidationException("Drought threshold year range is outside the output year range.")

    def configure_acc_water(self, acc_water_config):
        """ini_reader.py:473-486."""
        self.ResCapacityFile = os.path.join(self.AccWatDir, acc_water_config['ResCapacityFile'])
        self.BfiFile = os.path.join(self.AccWatDir, acc_water_config['BfiFile'])
        self.HistEndYear = int(acc_water_config['HistEndYear'])
        self.GCAM_StartYear = self.ck_year(int(acc_water_config['GCAM_StartYear']))
        self.GCAM_EndYear = int(acc_water_config['GCAM_EndYear'])
        self.GCAM_YearStep = int(acc_water_config['GCAM_YearStep'])
        self.MovingMeanWindow = int(acc_water_config['MovingMeanWindow'])
        self.Env_FlowPercent = float(acc_water_config['Env_FlowPercent'])
        if (self.StartYear > self.GCAM_StartYear) or (self.EndYear < self.GCAM_EndYear):
            raise ValidationException("Accessible water range of GCAM years are outside "
                                      "the range of years in climate data.")

    def configure_timeseries_plot(self, timeseries_config):
        """ini_reader.py:443-458."""
        self.TimeSeriesScale = int(timeseries_config['Scale'])
        self.TimeSeriesMapID = 999
        try:
            self.TimeSeriesMapID = int(timeseries_config['MapID'])
        except TypeError:
            self.TimeSeriesMapID = list(map(int, timeseries_config['MapID']))

    def ck_year(self, yr):
        """ini_reader.py:545-551."""
        if (yr < self.StartYear) or (yr > self.EndYear):
            raise ValidationException("Accessible water year {0} is outside the range of years in the climate "
                                      "data ({1}-{2}).".format(yr, self.StartYear, self.EndYear))
        return yr

    def configure_calibration(self, calibration_config):
        """ini_reader.py:504-519."""
        self.set_calibrate = int(calibration_config['set_calibrate'])
        self.cal_observed = calibration_config['observed']
        self.obs_unit = self.ck_obs_unit(self.set_calibrate, calibration_config['obs_unit'])
        self.calib_out_dir = self.create_dir(calibration_config['calib_out_dir'])
        try:
            self.cal_basins = calibration_config['calibration_basins']
            if type(self.cal_basins) is not list:
                self.cal_basins = [self.cal_basins]
        except KeyError:
            self.cal_basins = ['1-{}'.format(self.n_basins)]

    @staticmethod
    def ck_obs_unit(set_calib, unit):
        """ini_reader.py:521-545."""
        valid_runoff = ('km3_per_mth', 'mm_per_mth')
        valid_streamflow = ('m3_per_sec')
        if set_calib == 0:
            if unit not in valid_runoff:
                raise ValidationException("Calibration data input units '{}' for runoff data "
                                          "not in required units '{}'".format(unit, valid_runoff))
            return unit
        elif set_calib == 1:
            if unit not in valid_streamflow:
                raise ValidationException("Calibration data input units '{}' for streamflow data "
                                          "not in required units '{}'".format(unit, valid_streamflow))
            return unit

    def custom_runoff(self, f):
        return None if f == 'none' else os.path.join(self.rt_model_dir, f)

    @staticmethod
    def create_dir(pth):
        if os.path.isdir(pth) is False:
            os.makedirs(pth, exist_ok=True)
        return pth

    def log_info(self):
        logging.info('ProjectName : {}'.format(self.ProjectName))
        logging.info('InputFolder : {}'.format(self.InputFolder))
        logging.info('OutputFolder: {}'.format(self.OutputFolder))
        logging.info('StartYear - End Year: {0}-{1}'.format(self.StartYear, self.EndYear))
        logging.info('Number of Months    : {}'.format(self.nmonths))
        if str(self.HistFlag).lower() in ['true', 't', 'yes', 'y', '1']:
            logging.info('Running: Historic Mode')
        else:
            logging.info('Running: Future Mode')

    def update(self, args):
        """Overwrite configuration options (ini_reader.py:598-607)."""
        for k, v in args.items():
            if not hasattr(self, k):
                print('Warning: {} is not a valid parameter'.format(k))
            setattr(self, k, v)
