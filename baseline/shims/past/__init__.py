"""Stub of the `past` package of python-future (absent from this image and from the offline
wheelhouse).  The reference imports one name from it (`numbskull/numbskull.py:6`)."""
