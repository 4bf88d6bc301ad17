"""
Model interface - drop-in for xanthos/model.py (Xanthos, run_model; :21-132).

    from xanthos_b200 import Xanthos, run_model
    res = Xanthos('pm_abcd_mrtm.ini').execute(args={...})     # returns the Components object
    res.Q.shape == (ncell, nmonths)
"""

import argparse
import logging
import os
import sys

from .data_reader.ini_reader import ConfigReader
from .configurations import ConfigRunner


class Xanthos:
    """An extensible global hydrologic model (hot path on the B200)."""

    def __init__(self, ini):
        self.ini = ini
        self.config = None

    @staticmethod
    def make_dir(pth):
        if not os.path.exists(pth):
            os.makedirs(pth)

    def init_log(self):
        """Project-wide logger to stdout and <OutputFolder>/logfile.log (model.py:46-69)."""
        log_format = logging.Formatter('%(levelname)s: %(message)s')
        logger = logging.getLogger()
        logger.setLevel(logging.DEBUG)
        c_handler = logging.StreamHandler(sys.stdout)
        c_handler.setLevel(logging.DEBUG)
        c_handler.setFormatter(log_format)
        logger.addHandler(c_handler)
        f_handler = logging.FileHandler(os.path.join(self.config.OutputFolder, 'logfile.log'))
        f_handler.setFormatter(log_format)
        logger.addHandler(f_handler)

    def stage(self, mem_args):
        self.config = ConfigReader(self.ini)
        self.config.update(mem_args)
        self.make_dir(self.config.OutputFolder)
        self.init_log()

    def execute(self, args={}):
        """Run the configuration; `args` overrides config attributes, e.g. with in-memory arrays (model.py:82-98)."""
        self.stage(args)
        self.config.log_info()
        results = ConfigRunner(self.config).run()
        self.cleanup()
        return results

    def cleanup(self):
        logging.info("End of {0}".format(self.config.ProjectName))
        logger = logging.getLogger()
        for handler in logger.handlers[:]:
            handler.close()
            logger.removeHandler(handler)


def run_model(config_file):
    """Run the model based on a user-defined configuration file (model.py:111-122)."""
    Xanthos(config_file).execute()


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument('config_file', type=str, help='Full path with file name to INI configuration file.')
    a = parser.parse_args()
    Xanthos(a.config_file).execute()
