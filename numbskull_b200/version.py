"""Package version: the reference API level this drop-in mirrors (numbskull 0.1.1,
``numbskull/version.py``) plus a local tag for the B200 build."""
REFERENCE_API = (0, 1, 1)
__version__ = ".".join(str(x) for x in REFERENCE_API) + "+b200"
