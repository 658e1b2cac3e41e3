"""Locate the engine package from inside the drop-in ``lib/`` tree.

The reference puts ``lib/`` on sys.path (tools/_init_paths.py:19-22) and imports ``models``,
``core``, ``utils`` and ``config`` as top-level packages.  This tree is used the same way, so
the engine (vae-2_b200/engine) is loaded by file location under a fixed module name.
"""
import importlib.util
import os
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # .../vae-2_b200
_NAME = "vae2_b200_engine"


def engine():
    mod = sys.modules.get(_NAME)
    if mod is None:
        path = os.path.join(_PKG_ROOT, "engine")
        spec = importlib.util.spec_from_file_location(_NAME, os.path.join(path, "__init__.py"),
                                                      submodule_search_locations=[path])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[_NAME] = mod
        spec.loader.exec_module(mod)
    return mod
