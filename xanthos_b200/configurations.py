"""
Model configurations - drop-in for xanthos/configurations.py (ConfigRunner, :17-141).

Which of PET / runoff / routing run, which of them are step-wise (hargreaves, gwam) and whether the
whole-model spin-up pass runs (gwam) is decided exactly as in the reference (:54-87, :106-123).
"""

import logging

from .components import Components


class ConfigRunner:
    """Run the components specified by the configuration file."""

    def __init__(self, config):
        PET_COMPONENTS = ['hs', 'hargreaves', 'pm', 'thornthwaite']
        RUNOFF_COMPONENTS = ['abcd', 'gwam']
        ROUTING_COMPONENTS = ['mrtm']
        SPINUP_COMPONENTS = ['gwam']                 # runoff components that need the whole model to spin up
        PET_STEPWISE_COMPONENTS = ['hargreaves']
        RUNOFF_STEPWISE_COMPONENTS = ['gwam']
        ROUTING_STEPWISE_COMPONENTS = []
        self.run_pet = config.pet_module in PET_COMPONENTS
        self.run_runoff = config.runoff_module in RUNOFF_COMPONENTS
        self.run_routing = config.routing_module in ROUTING_COMPONENTS
        self.spinup = config.runoff_module in SPINUP_COMPONENTS
        self.pet_timestep = config.nmonths * (config.pet_module in PET_STEPWISE_COMPONENTS)
        self.runoff_timestep = config.nmonths * (config.runoff_module in RUNOFF_STEPWISE_COMPONENTS)
        self.routing_timestep = config.nmonths * (config.routing_module in ROUTING_STEPWISE_COMPONENTS)
        self.config = config

    def run(self):
        """Run all valid components; returns the Components object (configurations.py:89-141)."""
        if not (self.run_pet or self.run_runoff or self.run_routing):
            logging.warning("Selected configuration {0} not supported.".format(self.config.mod_cfg))
            return
        c = Components(self.config)
        if self.spinup:
            c.simulation(run_pet=self.run_pet, run_runoff=self.run_runoff, run_routing=self.run_routing,
                         pet_num_steps=self.config.runoff_spinup, runoff_num_steps=self.config.runoff_spinup,
                         routing_num_steps=self.config.routing_spinup, notify='Spin Up')
        c.simulation(run_pet=self.run_pet, run_runoff=self.run_runoff, run_routing=self.run_routing,
                     pet_num_steps=self.pet_timestep, runoff_num_steps=self.runoff_timestep,
                     routing_num_steps=self.routing_timestep, notify='Simulation')
        c.accessible_water()
        c.drought()
        c.hydropower_potential()
        c.hydropower_actual()
        c.diagnostics()
        if not self.config.calibrate:
            c.output_simulation()
        c.plots()
        return c
