"""CPU checker for the GPU Jones-Plassmann colouring (TEST INFRASTRUCTURE ONLY).

Jones-Plassmann with "smallest free colour" and fixed priorities is
schedule-independent: it equals a sequential greedy colouring that visits the
variables in decreasing priority.  This module computes that greedy colouring
with the same priority function as ``csrc/nb_build.cu`` (``nb_jp_priority``)
so the device result can be checked bit-exactly, and verifies validity (no two
same-colour variables share a factor).

``policy_coloring`` restates the library's choice between the two priority
orders (``nb_build.cu color_graph``): hashed priorities; if that needs more than
two colours, the natural order (smaller global id first) provided its round
count -- the longest chain of "smaller-id neighbour" dependencies -- stays under
the cap and it uses fewer colours."""
import numpy as np

M64 = (1 << 64) - 1


def _mix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def jp_priority(gid, seed):
    return (_mix64(gid ^ _mix64(seed)) & 0xFFFFFFFF00000000) | (gid & 0xFFFFFFFF)


def neighbours(variable, factor, fmap):
    nvar = len(variable)
    adj = [set() for _ in range(nvar)]
    owned = variable["isEvidence"] != 4
    for f in factor:
        mem = fmap["vid"][f["ftv_offset"]:f["ftv_offset"] + f["arity"]]
        mem = [int(m) for m in mem if owned[m]]
        for a in mem:
            for b in mem:
                if a != b:
                    adj[a].add(b)
    return adj


NATURAL_ROUND_CAP = 65536


def _greedy(variable, adj, gid, prio):
    nvar = len(variable)
    color = np.full(nvar, -1, np.int32)
    for v in sorted(range(nvar), key=lambda i: -prio[i]):
        if variable["isEvidence"][v] == 4:
            continue
        used = {int(color[u]) for u in adj[v] if color[u] >= 0}
        c = 0
        while c in used:
            c += 1
        color[v] = c
    return color


def natural_rounds(variable, adj, gid):
    """Rounds the natural-order colouring takes when a colour becomes visible one round after
    it is taken (k_jp_round mode 1): the longest chain of smaller-id neighbours.  (Fewer than 64
    colours assumed: no colour-window retries.)"""
    nvar = len(variable)
    level = np.zeros(nvar, np.int64)
    for v in sorted(range(nvar), key=lambda i: int(gid[i])):
        if variable["isEvidence"][v] == 4:
            continue
        level[v] = 1 + max((level[u] for u in adj[v] if gid[u] < gid[v]), default=0)
    return int(level.max()) if nvar else 0


def policy_coloring(variable, factor, fmap, seed, global_vid=None, cap=NATURAL_ROUND_CAP):
    """The colouring the library picks; returns (colours, mode) with mode 0 = hashed, 1 = natural."""
    nvar = len(variable)
    adj = neighbours(variable, factor, fmap)
    gid = np.arange(nvar) if global_vid is None else np.asarray(global_vid)
    hashed = _greedy(variable, adj, gid, [jp_priority(int(gid[v]), seed) for v in range(nvar)])
    pairwise = len(factor) == 0 or int(factor["arity"].max()) <= 2     # nb_build.cu color_graph: max arity <= 2
    if nvar and hashed.max() + 1 > 2 and cap > 0 and pairwise and natural_rounds(variable, adj, gid) <= cap:
        natural = _greedy(variable, adj, gid, [-int(gid[v]) for v in range(nvar)])
        if natural.max() < hashed.max():
            return relabel_by_min_id(natural, gid), 1
    return relabel_by_min_id(hashed, gid), 0


def greedy_coloring(variable, factor, fmap, seed, global_vid=None):
    """Hashed-priority colouring alone (what the library yields with the natural order disabled)."""
    nvar = len(variable)
    adj = neighbours(variable, factor, fmap)
    gid = np.arange(nvar) if global_vid is None else np.asarray(global_vid)
    prio = [jp_priority(int(gid[v]), seed) for v in range(nvar)]
    color = np.full(nvar, -1, np.int32)
    for v in sorted(range(nvar), key=lambda i: -prio[i]):
        if variable["isEvidence"][v] == 4:
            continue
        used = {int(color[u]) for u in adj[v] if color[u] >= 0}
        c = 0
        while c in used:
            c += 1
        color[v] = c
    return relabel_by_min_id(color, gid)


def relabel_by_min_id(color, gid):
    """Colours are numbered in increasing order of the smallest global id they contain
    (csrc/nb_build.cu order_colors_by_min_id)."""
    color = np.asarray(color).copy()
    n = int(color.max()) + 1 if len(color) else 0
    if n == 0:
        return color
    mins = np.full(n, np.iinfo(np.int64).max, np.int64)
    owned = color >= 0
    np.minimum.at(mins, color[owned], np.asarray(gid)[owned])
    order = np.argsort(mins, kind="stable")
    cmap = np.empty(n, np.int32)
    cmap[order] = np.arange(n, dtype=np.int32)
    color[owned] = cmap[color[owned]]
    return color


def conflicts(variable, factor, fmap, color):
    adj = neighbours(variable, factor, fmap)
    return sum(1 for v in range(len(variable)) for u in adj[v] if color[v] >= 0 and color[u] == color[v])
