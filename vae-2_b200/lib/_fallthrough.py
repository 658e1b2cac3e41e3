"""Fall-through from this drop-in ``lib/`` tree to a reference ``lib/`` tree that sits LATER on sys.path.

The mirror only re-implements the VAE^2 hot path.  Everything else the reference's own
drivers import -- ``core.function`` (epoch loops), ``datasets``, ``utils.modelsummary``, ``utils.metric``,
``models.seg_hrnet``, the legacy segmentation wrappers/losses inside ``utils.utils`` / ``core.criterion`` -- is the
reference's control plane and stays the reference's code:

* sub-MODULES the mirror does not have resolve through ``pkgutil.extend_path`` in each package's ``__init__``
  (e.g. ``core.function`` is found in <reference>/lib/core because <mirror>/lib/core has no function.py);
* NAMES the mirror's ``utils.utils`` / ``core.criterion`` do not define are looked up in the reference's module of the
  same name, loaded under a private module name (``reference_attr`` below, used from a PEP 562 ``__getattr__``).

With no reference tree on sys.path the mirror still works stand-alone; only those legacy names are then missing.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def reference_module(modname):
    """The reference's ``modname`` (e.g. 'utils.utils') from the first sys.path entry other than this tree that has
    it, loaded as ``_vae2_ref.<modname>``; None when no such tree is on sys.path."""
    if modname in _cache:
        return _cache[modname]
    rel = os.path.join(*modname.split(".")) + ".py"
    mod = None
    for p in sys.path:
        if not p or os.path.abspath(p) == _HERE:
            continue
        f = os.path.join(p, rel)
        if os.path.isfile(f):
            name = "_vae2_ref." + modname
            # keep relative imports of the reference file working: parent package = the (extended) mirror package
            spec = importlib.util.spec_from_file_location(name, f)
            mod = importlib.util.module_from_spec(spec)
            mod.__package__ = modname.rpartition(".")[0]
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            break
    _cache[modname] = mod
    return mod


def reference_attr(modname, name):
    ref = reference_module(modname)
    if ref is None or not hasattr(ref, name):
        raise AttributeError(
            "module %r of the vae2_b200 drop-in has no attribute %r, and no reference lib/ tree on sys.path "
            "provides it (legacy, off-path names fall through to the reference: put <reference>/lib AFTER this "
            "tree on sys.path)" % (modname, name))
    return getattr(ref, name)
