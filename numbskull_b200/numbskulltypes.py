"""Record layouts of the factor-graph arrays (host input ABI).

Same field names, order, widths and packing as the reference's
``numbskull/numbskulltypes.py:11-39`` so arrays built for numbskull can be
handed over unchanged.  The CUDA library parses these packed records directly
(``include/numbskull_b200.h``: ``nb_weight_rec`` ... ``nb_vtf_rec``); the
itemsize checks below keep the two sides honest.
"""
import numpy as np


def _rec(*fields):
    return np.dtype([(n, t) for n, t in fields])


Meta = _rec(("weights", np.int64), ("variables", np.int64),
            ("factors", np.int64), ("edges", np.int64))

Weight = _rec(("isFixed", np.bool_), ("initialValue", np.float64))

Variable = _rec(("isEvidence", np.int8), ("initialValue", np.int64),
                ("dataType", np.int16), ("cardinality", np.int64),
                ("vtf_offset", np.int64))

Factor = _rec(("factorFunction", np.int16), ("weightId", np.int64),
              ("featureValue", np.float64), ("arity", np.int64),
              ("ftv_offset", np.int64))

FactorToVar = _rec(("vid", np.int64), ("dense_equal_to", np.int64))

VarToFactor = _rec(("value", np.int64), ("factor_index_offset", np.int64),
                   ("factor_index_length", np.int64))

UnaryFactorOpt = _rec(("vid", np.int64), ("weightId", np.int64))

for _dt, _sz in ((Weight, 9), (Variable, 27), (Factor, 34), (FactorToVar, 16),
                 (VarToFactor, 24), (Meta, 32)):
    assert _dt.itemsize == _sz, (_dt, _sz)
