"""`from SDMplugin import *` -- the import line of the reference's scripts (example/test.py:11,
example/test_explicit.py:11; the SWIG module python/SDMplugin.i builds) resolves to the B200
implementation: LangevinIntegratorSDM with the setters / getters of SDMplugin.i:86-145, SDMUtils
with the constants of python/SDMUtils.py:9-15, OpenMMException.

All of it lives in openmm_sdm_plugin_b200/sdmplugin.py; this file only gives it the reference's
module name.  It needs the repository root on sys.path (or PYTHONPATH), like the reference needs its
build directory there."""
from openmm_sdm_plugin_b200.sdmplugin import *          # noqa: F401,F403
from openmm_sdm_plugin_b200.sdmplugin import LangevinIntegratorSDM, OpenMMException, SDMUtils   # noqa: F401

__all__ = ["LangevinIntegratorSDM", "OpenMMException", "SDMUtils"]
