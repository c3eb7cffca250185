"""Loaders and the variable-to-factor index builder (reference:
numbskull/dataloading.py).  Same function names and in-place semantics; the
work is done by the host C++ routines of libnumbskull_b200.so (nb_host.cpp)
instead of numba loops, bit-exact with the reference's outputs."""
from __future__ import print_function

import numpy as np

from . import _lib


def dataType(i):
    """dataloading.py:10-13."""
    return {0: "Boolean", 1: "Categorical"}.get(i, "Unknown")


def assign_vtf_offsets(variable):
    """The O(V) interpreter loop of numbskull.py:219-227 / :309-317; fills
    ``vtf_offset`` in place and returns the number of VarToFactor records."""
    import ctypes as C
    n = C.c_int64(0)
    _lib.check(_lib.lib().nb_assign_vtf_offsets(_lib.ptr(variable), len(variable), C.byref(n)))
    return int(n.value)


def compute_var_map(variables, factors, fmap, vmap, factor_index, domain_mask,
                    factors_to_skip=np.empty(0, np.int64)):
    """dataloading.py:16-81 (in place)."""
    dm = np.ascontiguousarray(domain_mask).view(np.uint8)
    fts = _lib.contiguous(factors_to_skip, np.int64)
    _lib.check(_lib.lib().nb_compute_var_map(
        _lib.ptr(variables), len(variables), _lib.ptr(factors), len(factors), _lib.ptr(fmap), len(fmap),
        _lib.ptr(vmap), len(vmap), _lib.ptr(factor_index), len(factor_index), _lib.ptr(dm),
        _lib.ptr(fts), len(fts)))


def _bytes(data):
    return np.ascontiguousarray(np.asarray(data).view(np.uint8))


def load_weights(data, nweights, weights):
    """dataloading.py:103-123."""
    data = _bytes(data)
    _lib.check(_lib.lib().nb_load_weights(_lib.ptr(data), data.size, int(nweights), _lib.ptr(weights)))
    print("LOADED WEIGHTS")


def load_variables(data, nvariables, variables):
    """dataloading.py:126-156."""
    data = _bytes(data)
    _lib.check(_lib.lib().nb_load_variables(_lib.ptr(data), data.size, int(nvariables), _lib.ptr(variables)))
    print("LOADED VARS")


def load_domains(data, domain_mask, vmap, variables):
    """dataloading.py:159-187."""
    data = _bytes(data)
    dm = domain_mask.view(np.uint8)
    _lib.check(_lib.lib().nb_load_domains(_lib.ptr(data), data.size, _lib.ptr(dm), _lib.ptr(vmap), len(vmap),
                                          _lib.ptr(variables), len(variables)))
    print("LOADED DOMAINS")


def load_factors(data, nfactors, factors, fmap, domain_mask, variable, vmap):
    """dataloading.py:190-237."""
    data = _bytes(data)
    dm = domain_mask.view(np.uint8)
    _lib.check(_lib.lib().nb_load_factors(_lib.ptr(data), data.size, int(nfactors), _lib.ptr(factors),
                                          _lib.ptr(fmap), len(fmap), _lib.ptr(dm), _lib.ptr(variable),
                                          len(variable), _lib.ptr(vmap), len(vmap)))
    print("LOADED FACTORS")
