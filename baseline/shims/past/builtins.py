"""`from past.builtins import long` -- the only name the reference needs (numbskull/numbskull.py:6)."""
long = int
