"""Drop-in alias: ``import numbskull`` resolves to the B200-native implementation.

The reference's own scripts (``test.py``: ``from numbskull import numbskull``;
``loadfg.py`` / ``test_lf_learning.py``: ``import numbskull``,
``from numbskull.numbskulltypes import *``, ``numbskull.inference.FACTORS``) run
unmodified against this package.  Everything lives in ``numbskull_b200``.
"""
import sys as _sys

import numbskull_b200 as _impl
from numbskull_b200 import NumbSkull, main, load, __version__  # noqa: F401
from numbskull_b200 import numbskull, numbskulltypes, inference, factorgraph, dataloading  # noqa: F401

for _name in ("numbskull", "numbskulltypes", "inference", "factorgraph", "dataloading", "timer", "version"):
    _sys.modules[__name__ + "." + _name] = getattr(_impl, _name, None) or __import__(
        "numbskull_b200." + _name, fromlist=[_name])

__all__ = ('numbskull', 'NumbSkull', 'main')
