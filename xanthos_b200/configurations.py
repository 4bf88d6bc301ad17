"""
Model configurations - drop-in for xanthos/configurations.py (ConfigRunner, :17-141).

Which of PET / runoff / routing run is decided exactly as in the reference (:70-87).  The
step-wise components of the reference (hargreaves, gwam) are not part of this build, so every
module iterates internally (time-step arguments are always 0).
"""

import logging

from .components import Components


class ConfigRunner:
    """Run the components specified by the configuration file."""

    def __init__(self, config):
        PET_COMPONENTS = ['hs', 'pm', 'thornthwaite']
        RUNOFF_COMPONENTS = ['abcd']
        ROUTING_COMPONENTS = ['mrtm']
        self.run_pet = config.pet_module in PET_COMPONENTS
        self.run_runoff = config.runoff_module in RUNOFF_COMPONENTS
        self.run_routing = config.routing_module in ROUTING_COMPONENTS
        self.spinup = False          # only gwam needs a whole-model spin-up pass (configurations.py:75)
        self.pet_timestep = 0
        self.runoff_timestep = 0
        self.routing_timestep = 0
        self.config = config

    def run(self):
        """Run all valid components; returns the Components object (configurations.py:89-141)."""
        if not (self.run_pet or self.run_runoff or self.run_routing):
            logging.warning("Selected configuration {0} not supported.".format(self.config.mod_cfg))
            return
        c = Components(self.config)
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
