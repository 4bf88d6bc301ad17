"""
Model interface - drop-in for xanthos/model.py (Xanthos, run_model; :21-132).

    from xanthos_b200 import Xanthos, run_model
    res = Xanthos('pm_abcd_mrtm.ini').execute(args={...})     # returns the Components object
    res.Q.shape == (ncell, nmonths)

Same public surface as the reference class (`make_dir`, `init_log`, `stage`, `execute`, `cleanup`) and the same
side effects: the output folder is created, every log record goes to stdout and to <OutputFolder>/logfile.log, and
the handlers are detached again when the run ends.
"""

import argparse
import logging
import os
import sys

from .configurations import ConfigRunner
from .data_reader.ini_reader import ConfigReader

LOG_FORMAT = '%(levelname)s: %(message)s'


class Xanthos:
    """An extensible global hydrologic model (hot path on the B200)."""

    def __init__(self, ini):
        self.ini = ini
        self.config = None
        self._handlers = []

    @staticmethod
    def make_dir(pth):
        os.makedirs(pth, exist_ok=True)

    def init_log(self):
        """Root logger at DEBUG with a console and a file handler (model.py:46-69)."""
        root = logging.getLogger()
        root.setLevel(logging.DEBUG)
        targets = (logging.StreamHandler(sys.stdout),
                   logging.FileHandler(os.path.join(self.config.OutputFolder, 'logfile.log')))
        for handler in targets:
            handler.setLevel(logging.DEBUG)
            handler.setFormatter(logging.Formatter(LOG_FORMAT))
            root.addHandler(handler)
            self._handlers.append(handler)

    def stage(self, mem_args):
        """Parse the .ini, apply the in-memory overrides, create the output folder, start logging (model.py:71-80)."""
        self.config = ConfigReader(self.ini)
        self.config.update(mem_args)
        self.make_dir(self.config.OutputFolder)
        self.init_log()

    def execute(self, args=None):
        """Run the configuration and return the Components object; `args` overrides config attributes, e.g. forcing
        passed as in-memory arrays instead of file names (model.py:82-98)."""
        self.stage(args or {})
        try:
            self.config.log_info()
            return ConfigRunner(self.config).run()
        finally:
            self.cleanup()

    def cleanup(self):
        """Close and detach every handler of the root logger, as the reference does (model.py:100-108)."""
        logging.info("End of {0}".format(self.config.ProjectName))
        root = logging.getLogger()
        for handler in list(root.handlers):
            handler.close()
            root.removeHandler(handler)
        self._handlers.clear()


def run_model(config_file):
    """Run the model based on a user-defined configuration file (model.py:111-122)."""
    Xanthos(config_file).execute()


if __name__ == "__main__":
    cli = argparse.ArgumentParser()
    cli.add_argument('config_file', type=str, help='Full path with file name to INI configuration file.')
    Xanthos(cli.parse_args().config_file).execute()
