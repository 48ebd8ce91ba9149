"""Package-level switches (counterpart of ``glimslib/config.py:5-23``)."""
import os

USE_ADJOINT = False                      # the discrete adjoint is a "next" row
base_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
output_dir = os.path.join(base_path, "output")
