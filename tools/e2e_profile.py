import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
import numbskull_b200 as nb
from numbskull_b200 import _lib, synth
ns = nb.NumbSkull(quiet=True); ns.loadFactorGraph(*synth.ising_grid(4096, 4096)); fg = ns.factorGraphs[0]; fg.seed = 1
L = _lib.lib(); g = fg._device_graph()
fg.inference(0, 1, sample_evidence=True)
def t(fn, n=5):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
vv = fg.var_value[0]; cnt = fg.count; mg = fg.marginals; w = fg.weight_value[0]
print("set_var_values  %.2f ms" % t(lambda: _lib.check(L.nb_set_var_values(g, 0, _lib.ptr(vv)))))
print("set_weights     %.2f ms" % t(lambda: _lib.check(L.nb_set_weights(g, _lib.ptr(w)))))
print("reset_counts    %.2f ms" % t(lambda: (_lib.check(L.nb_reset_counts(g)), _lib.check(L.nb_synchronize(g)))))
print("sweep           %.2f ms" % t(lambda: _lib.check(L.nb_gibbs_sweeps(g, 1, 0, 1, 1))))
print("get_var_values  %.2f ms" % t(lambda: _lib.check(L.nb_get_var_values(g, 0, _lib.ptr(vv)))))
print("get_counts_marg %.2f ms" % t(lambda: _lib.check(L.nb_get_counts_marginals(g, _lib.ptr(cnt), 1, _lib.ptr(mg), 1.0))))
print("inference(0,1)  %.2f ms" % t(lambda: fg.inference(0, 1, sample_evidence=True)))
