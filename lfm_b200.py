"""Import shim: the package directory is named ``lightfieldmicroscopy_pc-bzip2_b200`` (with a hyphen), which the
``import`` statement cannot spell. ``import lfm_b200`` gives the same module."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_mod = importlib.import_module("lightfieldmicroscopy_pc-bzip2_b200")
sys.modules[__name__] = _mod
