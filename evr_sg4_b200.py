"""Importable alias of the hyphen-named package ``elvibrot-tnumtana_b200``."""
import importlib as _il
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
_pkg = _il.import_module("elvibrot-tnumtana_b200")
_sys.modules[__name__] = _pkg
