"""
refshim.py -- import the UNMODIFIED reference (michael-petersen/exptool) read-only.

TEST INFRASTRUCTURE ONLY.  Used (a) by tests/golden/make_golden.py to generate
the committed golden vectors and (b) by tests that cross-check the oracle
against the live reference when /root/reference is present (this container;
never the GPU box).  Nothing in exptool_b200/ imports this.

The reference imports matplotlib / skimage at module import time
(eof.py:66-69, utils/utils.py:62, orbits/orbit.py:33-35); neither is installed
here, so they are replaced by MagicMock before import.
"""
import os
import sys
import importlib
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get('EXPTOOL_REFERENCE', '/root/reference')

_STUBS = ['matplotlib', 'matplotlib.pyplot', 'matplotlib.cm', 'matplotlib.colors',
          'matplotlib.colorbar', 'matplotlib.patches', 'matplotlib.ticker',
          'matplotlib.gridspec', 'matplotlib.collections',
          'mpl_toolkits', 'mpl_toolkits.mplot3d', 'mpl_toolkits.axes_grid1',
          'skimage', 'skimage.measure']


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'exptool'))


def load():
    """Return a dict of the reference modules on the hot path."""
    if not available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = MagicMock()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mods = {}
    for short, full in [('eof', 'exptool.basis.eof'),
                        ('spheresl', 'exptool.basis.spheresl'),
                        ('potential', 'exptool.basis.potential'),
                        ('compatibility', 'exptool.basis.compatibility'),
                        ('halo_methods', 'exptool.utils.halo_methods'),
                        ('integrate', 'exptool.utils.integrate'),
                        ('particle', 'exptool.io.particle'),
                        ('orbit', 'exptool.orbits.orbit'),
                        ('pattern', 'exptool.analysis.pattern')]:
        mods[short] = importlib.import_module(full)
    return mods
