"""Measure the per-colour cost of the peer-to-peer halo exchange (run under torchrun, 2+ ranks)."""
import ctypes as C, datetime, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from numbskull_b200 import _lib, partition

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
run = partition.ising_strip_runner(1024, 4096, rank, world, local, seed=1)
L, g = _lib.lib(), run.fg._g
run.fg._upload(0, 0, evid=False)
run.sweeps(3, True, True)
torch.cuda.synchronize(); dist.barrier()
for mode, mask in (("blocking", 1), ("nowait", 1 | 16)):
    n = 2000
    torch.cuda.synchronize(); dist.barrier()
    _lib.check(L.nb_timer_start(g))
    for i in range(n):
        _lib.check(L.nb_p2p_exchange(g, (2 * i) % run.n_phases if run.split else i % run.n_phases, mask))
    _lib.check(L.nb_p2p_wait(g))
    ms = C.c_float(0); _lib.check(L.nb_timer_stop(g, C.byref(ms)))
    if rank == 0:
        print("%s: %.2f us per exchange (%d colours, %d boundary values per sweep)" % (mode, 1e3 * ms.value / n, run.n_colors, run.halo_bytes_per_sweep))
_lib.check(L.nb_p2p_check(g))
dist.barrier(); dist.destroy_process_group()
