"""Counterpart of ``glimslib/simulation/config.py``."""
import os
import tempfile

from glimslib_b200.config import USE_ADJOINT  # noqa: F401

output_dir_simulation_tmp = os.path.join(tempfile.gettempdir(), "glimslib_b200_simulation_tmp")
