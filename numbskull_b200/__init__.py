"""numbskull_b200: B200-native Gibbs sampling / weight learning behind
numbskull's Python API (``import numbskull_b200 as numbskull``)."""
from .numbskull import NumbSkull, main, load  # noqa: F401
from .version import __version__  # noqa: F401
from . import numbskull, numbskulltypes, inference, factorgraph, dataloading  # noqa: F401

__all__ = ('numbskull', 'NumbSkull', 'main')
