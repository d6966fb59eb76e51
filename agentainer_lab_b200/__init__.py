"""Importable alias of the `agentainer-lab_b200/` package directory (a hyphen cannot appear in an import
statement).  All code lives in `agentainer-lab_b200/`; this module only redirects the package search path."""
import os as _os

__path__ = [_os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), _os.pardir, "agentainer-lab_b200"))]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
