#!/usr/bin/env python
"""One profiled learning epoch of the labelling-function model (config 3 shape): run under
   ncu --profile-from-start off --set full --import-source on -k regex:k_learn_cells -c 1 ...
usage: prof_learn.py [copies] [n_lf]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import numbskull_b200 as nb  # noqa: E402
from numbskull_b200 import _lib, synth  # noqa: E402

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
n_lf = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ns = nb.NumbSkull(quiet=True)
ns.loadFactorGraph(*synth.lf_model(copies, n_lf, np.random.default_rng(1003)))
fg = ns.factorGraphs[0]
fg.seed = 1
g = fg._device_graph()
L = _lib.lib()
fg._sync_device(0, 0)


def epochs(n):
    s = C.c_double(1e-4)
    _lib.check(L.nb_learn_sweeps(g, n, C.byref(s), 1.0, 1, 0.01, 1.0, 1, fg.seed, 0))


epochs(1)
nblk = C.c_int(0)
_lib.check(L.nb_learn_blocks(g, 1e-4, 1, 0, C.byref(nblk)))
_lib.check(L.nb_timer_start(g))
epochs(3)
ms = C.c_float(0)
_lib.check(L.nb_timer_stop(g, C.byref(ms)))
info = fg.device_info()
print("copies %d: %.3f ms / epoch, %d blocks x %d colours = %d cells, %.2f us / cell"
      % (copies, ms.value / 3, nblk.value, info["n_colors"], nblk.value * info["n_colors"],
         1e3 * ms.value / 3 / (nblk.value * info["n_colors"])))
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
epochs(1)
rt.cudaProfilerStop()
