"""glimslib_b200 -- B200-native forward-simulation hot path of GlimSLib.

``glimslib_b200.engine.Engine`` is the ctypes handle on the CUDA library (``include/glims_b200.h``);
``glimslib_b200.simulation`` mirrors the reference's ``glimslib.simulation`` API on top of it.
"""
__version__ = "0.1.0"
