"""Path settings and package-level switches (counterpart of ``glimslib/config.py:5-23``; same names, so that
``from glimslib.config import *`` in the reference's scripts and ``testing_config`` modules keeps working).

``GLIMSLIB_OUTPUT_DIR`` overrides the output root (the reference hard-wires ``<repo>/output``)."""
import os

base_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

output_dir = os.environ.get("GLIMSLIB_OUTPUT_DIR", os.path.join(base_path, "output"))
output_dir_testing = os.path.join(output_dir, "test_cases")
output_dir_simulation = os.path.join(output_dir, "simulation")
output_dir_application = os.path.join(output_dir, "application")

output_dir_temp = os.path.join(output_dir, "temp")

test_dir = os.path.join(base_path, "test_cases")
test_data_dir = os.path.join(test_dir, "data")

# meshtool settings (names kept; MeshTool itself is outside the hot path)
path_to_meshtool = "/home/fenics/software/MESHTOOL_source"
path_to_meshtool_bin = os.path.join(path_to_meshtool, "bin", "MeshTool")
path_to_meshtool_xsd = os.path.join(path_to_meshtool, "src", "xml-io", "imaging_meshing_schema.xsd")

# Switch for using adjoint; false by default.  The dolfin-adjoint tape is not part of the B200 hot path.
USE_ADJOINT = False
