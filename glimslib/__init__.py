"""``glimslib`` -- import alias of :mod:`glimslib_b200`, so that scripts written against the reference
(``from glimslib.simulation.simulation_tumor_growth import TumorGrowth``, ``from glimslib import fenics_local as fenics``,
``import glimslib.utils.data_io as dio`` ...) run unchanged on the B200 backend.

Every ``glimslib.<x>`` import resolves to the *same module object* as ``glimslib_b200.<x>`` (no second copy of any class),
through a meta-path finder registered below.  Put the repository root on ``PYTHONPATH`` ahead of the reference tree.
"""
import importlib
import importlib.abc
import importlib.util
import sys

import glimslib_b200 as _impl

_PREFIX, _REAL = "glimslib", "glimslib_b200"


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_PREFIX + "."):
            return None
        real = _REAL + fullname[len(_PREFIX):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        return importlib.util.spec_from_loader(fullname, self, is_package=self._is_pkg(real))

    @staticmethod
    def _is_pkg(real):
        spec = importlib.util.find_spec(real)
        return spec is not None and spec.submodule_search_locations is not None

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_PREFIX):])

    def exec_module(self, module):          # the real module is already initialised
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

__version__ = _impl.__version__
__path__ = []          # no files of its own: sub-modules come from the finder above
