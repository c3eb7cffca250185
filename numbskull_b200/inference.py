"""Factor-function table of the hot path.

``FACTORS`` / ``FUNC_*`` carry the same names and ids as the reference
(``numbskull/inference.py:74-146``).  The functions themselves live in
``csrc/nb_eval.cuh`` (``nb_eval_incidence``); there is no Python or numba
implementation in this package.
"""
from .synth import FUNC as _FUNC

FACTORS = dict(_FUNC)
for _name, _value in FACTORS.items():
    globals()["FUNC_" + _name] = _value
del _name, _value
