#!/usr/bin/env python
"""One profiled sweep of the KBC (config 4) shape: run under
   ncu --profile-from-start off --set full --import-source on -k regex:k_gibbs_tt -c N ...
usage: prof_c4.py [nvar] [sweeps]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import numbskull_b200 as nb  # noqa: E402
from numbskull_b200 import _lib, synth  # noqa: E402

nvar = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ns = nb.NumbSkull(quiet=True)
ns.loadFactorGraph(*synth.kbc_fast(nvar, seed=1004))
fg = ns.factorGraphs[0]
fg.seed = 1
g = fg._device_graph()
L = _lib.lib()
fg._sync_device(0, 0)
_lib.check(L.nb_gibbs_sweeps(g, 3, 1, 1, fg.seed))
_lib.check(L.nb_timer_start(g))
_lib.check(L.nb_gibbs_sweeps(g, 5, 0, 1, fg.seed))
ms = C.c_float(0)
_lib.check(L.nb_timer_stop(g, C.byref(ms)))
print("nvar %d: %.3f ms / sweep, info %s" % (nvar, ms.value / 5, fg.device_info()))
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
_lib.check(L.nb_gibbs_sweeps(g, sweeps, 0, 1, fg.seed))
rt.cudaProfilerStop()
