"""Host logic of the multi-GPU path on CPU: block partitioning, ghost sets, the
halo plan built over torch.distributed (gloo, world_size 2) and the
gather -> point-to-point -> scatter exchange."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_extract_local_covers_the_graph():
    from numbskull_b200 import partition, synth
    w, v, f, fm, dm, e = synth.random_graph(60, 150, np.random.default_rng(3), max_arity=3, categorical_frac=0.3,
                                            funcs=(0, 1, 2, 3, 12), card=3)
    world = 3
    b = partition.block_bounds(len(v), world)
    seen_factors = np.zeros(len(f), int)
    for r in range(world):
        loc = partition.extract_local(w, v, f, fm, int(b[r]), int(b[r + 1]))
        gv, n_owned = loc["global_vid"], loc["n_owned"]
        assert np.array_equal(gv[:n_owned], np.arange(b[r], b[r + 1]))
        assert (loc["variable"]["isEvidence"][n_owned:] == 4).all()
        assert np.array_equal(loc["variable"]["cardinality"], v["cardinality"][gv])
        # every local factor maps back to a global factor with the same members, in global order
        gl_members = [tuple(fm["vid"][x["ftv_offset"]:x["ftv_offset"] + x["arity"]]) for x in f]
        keep = [i for i, m in enumerate(gl_members) if any(b[r] <= u < b[r + 1] for u in m)]
        assert len(keep) == len(loc["factor"])
        for li, gi in enumerate(keep):
            lf = loc["factor"][li]
            lm = tuple(gv[loc["fmap"]["vid"][lf["ftv_offset"]:lf["ftv_offset"] + lf["arity"]]])
            assert lm == gl_members[gi] and lf["weightId"] == f["weightId"][gi]
            seen_factors[gi] += 1
    assert (seen_factors[f["arity"] > 0] >= 1).all()


def test_interior_variables_never_touch_a_ghost():
    """The invariant behind the boundary / interior phases (nb_split_colors): an owned variable
    that shares a factor with a ghost is itself a ghost on the ghost's owner, i.e. it is in the
    set of variables this rank sends -- so everything outside that set only reads owned values."""
    from numbskull_b200 import partition, synth
    w, v, f, fm, dm, e = synth.random_graph(90, 260, np.random.default_rng(11), max_arity=4, categorical_frac=0.2,
                                            funcs=(0, 1, 2, 3, 12), card=3)
    world = 3
    b = partition.block_bounds(len(v), world)
    owner = np.searchsorted(b, np.arange(len(v)), side="right") - 1
    locs = [partition.extract_local(w, v, f, fm, int(b[r]), int(b[r + 1])) for r in range(world)]
    for r, loc in enumerate(locs):
        gv, n_owned = loc["global_vid"], loc["n_owned"]
        # what the other ranks hold as ghosts of r = what r must send (HaloPlan.send_ids, without the collective)
        sent = set()
        for q, other in enumerate(locs):
            if q != r:
                ghosts = other["global_vid"][other["n_owned"]:]
                sent.update(int(x) for x in ghosts[owner[ghosts] == r])
        n_boundary_checked = 0
        for lf in loc["factor"]:
            mem = loc["fmap"]["vid"][lf["ftv_offset"]:lf["ftv_offset"] + lf["arity"]]
            if (mem >= n_owned).any():                      # the factor has a ghost member
                for m in mem[mem < n_owned]:
                    assert int(gv[m]) in sent
                    n_boundary_checked += 1
        assert n_boundary_checked > 0


def test_ising_strip_matches_global_partition():
    from numbskull_b200 import partition, synth
    rows, cols, world = 3, 5, 3
    w, v, f, fm, dm, e = synth.ising_grid(rows * world, cols)
    b = partition.block_bounds(len(v), world)
    for r in range(world):
        ref = partition.extract_local(w, v, f, fm, int(b[r]), int(b[r + 1]))
        loc, n_global = partition.ising_strip(rows, cols, r, world)
        assert n_global == len(v)
        assert np.array_equal(loc["global_vid"], ref["global_vid"])
        assert np.array_equal(loc["fmap"], ref["fmap"]) and np.array_equal(loc["factor"], ref["factor"])
        assert np.array_equal(loc["variable"], ref["variable"])


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from numbskull_b200 import partition, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, v, f, fm, dm, e = synth.random_graph(200, 500, np.random.default_rng(5), max_arity=3)
    b = partition.block_bounds(len(v), world)
    loc = partition.extract_local(w, v, f, fm, int(b[rank]), int(b[rank + 1]))
    gv, n_owned = loc["global_vid"], loc["n_owned"]
    plan = partition.HaloPlan(gv, n_owned, lambda gids: np.searchsorted(b, gids, side="right") - 1, rank, world)
    # state: owners know f(global id); ghosts start at -1 and must be filled by the exchange
    x = torch.full((len(gv),), -1, dtype=torch.int32)
    x[:n_owned] = torch.from_numpy((gv[:n_owned] * 7 + 3).astype(np.int32))
    ex = partition.Exchange(plan.send_ids, plan.recv_ids, rank, world, torch.int32, "cpu")

    def gather(ids, out):
        out.copy_(x[ids.long()])

    def scatter(ids, buf):
        x[ids.long()] = buf

    ex.run(gather, scatter)
    ok = bool((x.numpy() == (gv * 7 + 3)).all())
    # a restricted (per-colour style) exchange only touches the flagged variables
    y = torch.full((len(gv),), -1, dtype=torch.int32)
    y[:n_owned] = x[:n_owned]
    flag = (gv % 3 == 0)
    send, recv = plan.restrict(flag)
    ex2 = partition.Exchange(send, recv, rank, world, torch.int32, "cpu")
    ex2.run(lambda ids, out: out.copy_(y[ids.long()]), lambda ids, buf: y.__setitem__(ids.long(), buf))
    want = np.where(flag | (np.arange(len(gv)) < n_owned), gv * 7 + 3, -1)
    ok2 = bool((y.numpy() == want).all())
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([ok, ok2, ex.n_send, ex.n_recv]))
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(str(tmp_path / ("r%d.npy" % r))) for r in range(world)]
    for r in res:
        assert r[0] == 1 and r[1] == 1
    assert res[0][2] == res[1][3] and res[0][3] == res[1][2] and res[0][2] > 0


def test_kbc_block_generator_matches_extract_local():
    """Partitioned runs of BASELINE config 4 generate each owner block on its own
    (synth.kbc_block, host threads, counter-based randomness): identical to cutting the block out
    of the whole graph (partition.extract_local on synth.kbc_fast)."""
    import numpy as np
    from numbskull_b200 import partition, synth
    n = 60_000
    w, v, f, fm, dm, e = synth.kbc_fast(n, seed=7)
    for world in (1, 3):
        for r in range(world):
            lo, hi = n * r // world, n * (r + 1) // world
            a = partition.extract_local(w, v, f, fm, lo, hi)
            b = synth.kbc_block(n, lo, hi, seed=7)
            for k in ("variable", "factor", "fmap", "global_vid", "weight"):
                assert np.array_equal(a[k], b[k]), (world, r, k)
            assert a["n_owned"] == b["n_owned"]
    # the whole-graph generator does not depend on the number of host threads
    import os
    assert len(f) == n + 3 * (n // 2) and int(f["arity"].sum()) == len(fm) == e


def test_owner_array_placement_matches_block_partition():
    """extract_local_by_owner with the block owners is extract_local; with a scrambled placement the
    ranks' shares still tile the graph (every variable owned once, every factor kept wherever it has
    an owned member, ghosts flagged isEvidence = 4)."""
    import numpy as np
    from numbskull_b200 import partition, synth
    w, v, f, fm, dm, e = synth.random_graph(500, 1200, np.random.default_rng(3), max_arity=3)
    world = 3
    bounds = partition.block_bounds(len(v), world)
    owner = (np.searchsorted(bounds, np.arange(len(v)), side="right") - 1).astype(np.int32)
    for r in range(world):
        a = partition.extract_local(w, v, f, fm, int(bounds[r]), int(bounds[r + 1]))
        b = partition.extract_local_by_owner(w, v, f, fm, owner, r)
        for k in ("variable", "factor", "fmap", "global_vid"):
            assert np.array_equal(a[k], b[k]), (r, k)
    owner = np.random.default_rng(4).integers(0, world, len(v)).astype(np.int32)
    seen = np.zeros(len(v), int)
    for r in range(world):
        loc = partition.extract_local_by_owner(w, v, f, fm, owner, r)
        n = loc["n_owned"]
        gv = loc["global_vid"]
        seen[gv[:n]] += 1
        assert (owner[gv[:n]] == r).all() and (owner[gv[n:]] != r).all()
        assert (loc["variable"]["isEvidence"][n:] == 4).all()
        assert (np.diff(gv[:n]) > 0).all() and (np.diff(gv[n:]) > 0).all()
        # every kept factor has an owned member, members translate back to the global ids
        ar = loc["factor"]["arity"].astype(int)
        first = loc["factor"]["ftv_offset"].astype(int)
        for i in range(0, len(ar), 37):
            m = loc["fmap"]["vid"][first[i]:first[i] + ar[i]]
            assert (m < n).any()
    assert (seen == 1).all()


def test_salt_partition_keys_and_locality_owners():
    """Reference partition keys (salt/src/messages.py:175-179; master keeps 'A' / 'B', minion i the
    keys that carry its id) and the locality-aware placement: on a windowed graph whose ids were
    shuffled, the RCM placement needs far fewer ghosts than blocks of the shuffled ids."""
    import numpy as np
    from numbskull_b200 import partition, synth
    own = partition.owners_from_salt_keys(["A", "B", "C0", "D0", "C1", "Du1", b"C2", "Au"], world=3)
    assert own.tolist() == [0, 0, 1, 1, 2, 2, 1, 0]
    assert partition.owners_from_salt_keys(["C5", "A"], world=1).tolist() == [0, 0]
    n = 20000
    w, v, f, fm, dm, e = synth.kbc_fast(n, seed=3, far_frac=0.0, hub_frac=0.0)
    perm = np.random.default_rng(1).permutation(n)
    fm2 = fm.copy()
    fm2["vid"] = perm[fm["vid"]]                    # same graph, ids carry no locality any more
    v2 = np.empty_like(v)
    v2[perm] = v
    world = 4
    bounds = partition.block_bounds(n, world)
    blocks = (np.searchsorted(bounds, np.arange(n), side="right") - 1).astype(np.int32)
    smart = partition.locality_owners(v2, f, fm2, world)
    assert np.bincount(smart, minlength=world).max() <= n // world + 1
    g_block = np.mean([partition.ghost_fraction(partition.extract_local_by_owner(w, v2, f, fm2, blocks, r)) for r in range(world)])
    g_smart = np.mean([partition.ghost_fraction(partition.extract_local_by_owner(w, v2, f, fm2, smart, r)) for r in range(world)])
    assert g_smart < 0.25 * g_block, (g_smart, g_block)


def test_block_ghosts_matches_numpy_and_rejects_bad_input():
    """nb_block_ghosts (threaded bitmap + prefix counts) against numpy's unique / searchsorted, the
    rule of partition.extract_local: ghosts ascending by global id, owned v -> v - lo, ghost -> n_owned
    + rank; empty member lists, empty blocks and blocks at either end; ids outside the graph are an
    error, not a crash."""
    import ctypes as C
    import numpy as np
    from numbskull_b200 import _lib
    from numbskull_b200.numbskulltypes import FactorToVar
    L = _lib.lib()
    rng = np.random.default_rng(5)
    nvar = 100_003
    for lo, hi, n in ((0, 40_000, 300_000), (60_000, nvar, 300_000), (50_000, 50_000, 1000), (10, 20, 0), (0, nvar, 5000)):
        fm = np.zeros(n, FactorToVar)
        fm["vid"] = rng.integers(0, nvar, n)
        fm["dense_equal_to"] = rng.integers(0, 7, n)
        gv = fm["vid"].copy()
        owned = (gv >= lo) & (gv < hi)
        want_ghosts = np.unique(gv[~owned])
        want_local = np.where(owned, gv - lo, (hi - lo) + np.searchsorted(want_ghosts, gv))
        ng = C.c_int64(-1)
        _lib.check(L.nb_block_ghosts(_lib.ptr(fm), n, nvar, lo, hi, None, C.byref(ng), 0))
        assert ng.value == len(want_ghosts)
        assert np.array_equal(fm["vid"], gv)                       # the counting call rewrites nothing
        ghosts = np.empty(ng.value, np.int64)
        _lib.check(L.nb_block_ghosts(_lib.ptr(fm), n, nvar, lo, hi, _lib.ptr(ghosts), C.byref(ng), 1))
        assert np.array_equal(ghosts, want_ghosts)
        assert np.array_equal(fm["vid"], want_local)
        assert np.array_equal(fm["dense_equal_to"], fm["dense_equal_to"])
    fm = np.zeros(4, FactorToVar)
    fm["vid"] = [1, 2, nvar, 3]
    ng = C.c_int64(0)
    with pytest.raises(Exception):
        _lib.check(L.nb_block_ghosts(_lib.ptr(fm), 4, nvar, 0, 2, None, C.byref(ng), 0))
    with pytest.raises(Exception):
        _lib.check(L.nb_block_ghosts(_lib.ptr(fm), 4, nvar, 5, 2, None, C.byref(ng), 0))


def test_extract_local_by_owner_threads_match_numpy():
    """nb_extract_local (host threads, bitmaps + prefix popcounts) against the numpy statement of the
    same rule, on a mixed graph under block, scrambled and lopsided placements (a rank that owns
    nothing included); broken factor records are an error."""
    import ctypes as C
    import numpy as np
    from numbskull_b200 import _lib, partition, synth
    w, v, f, fm, dm, e = synth.random_graph(3000, 8000, np.random.default_rng(21), max_arity=4, categorical_frac=0.2,
                                            funcs=(0, 1, 2, 3, 12), card=3)
    n = len(v)
    rng = np.random.default_rng(2)
    placements = [(np.arange(n) * 3 // n), rng.integers(0, 3, n), np.where(np.arange(n) < 10, 0, 1), np.zeros(n, int)]
    for owner in placements:
        owner = owner.astype(np.int32)
        for rank in range(3):
            a = partition.extract_local_by_owner(w, v, f, fm, owner, rank)
            b = partition._extract_local_by_owner_numpy(w, v, f, fm, owner, rank)
            for k in ("variable", "factor", "fmap", "global_vid", "domain_mask"):
                assert np.array_equal(a[k], b[k]), (rank, k)
            assert a["n_owned"] == b["n_owned"] == int((owner == rank).sum())
    bad = f.copy()
    bad["ftv_offset"][5] = len(fm)                 # members beyond fmap
    L = _lib.lib()
    z = [C.c_int64(0) for _ in range(4)]
    owner = np.zeros(n, np.int32)
    with pytest.raises(Exception):
        _lib.check(L.nb_extract_local(_lib.ptr(bad), len(bad), _lib.ptr(fm), len(fm), _lib.ptr(owner), n, 0,
                                      C.byref(z[0]), C.byref(z[1]), C.byref(z[2]), C.byref(z[3]), None, None, None))
