"""Importable alias of the package directory ``k3s-nvidia_b200/`` (a hyphen is not a valid module
name).  ``import k3s_nvidia_b200`` executes ``k3s-nvidia_b200/__init__.py`` with this module's
``__path__`` pointing there, so ``k3s_nvidia_b200.probe`` etc. resolve to the real files."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "k3s-nvidia_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
