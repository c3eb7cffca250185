"""Device build time of the C2 graph (colouring included) -- run on a GPU box."""
import sys, time
sys.path.insert(0, '/root/repo')
import numbskull_b200 as nb
from numbskull_b200 import synth
ns = nb.NumbSkull(quiet=True)
ns.loadFactorGraph(*synth.ising_grid(4096, 4096))
fg = ns.factorGraphs[0]
t0 = time.perf_counter()
fg._device_graph()
dt = time.perf_counter() - t0
info = fg.device_info()
print("device build %.2f s, %d colours, %d colouring rounds" % (dt, info["n_colors"], info["jp_rounds"]))
