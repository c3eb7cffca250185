"""CPU-side checks: the C-ABI library loads and exports what the header
declares, the host integer work (loaders, vtf offsets, compute_var_map) is
bit-exact against arrays produced by the real reference, and the Python
drop-in surface matches the reference's."""
import os
import re

import numpy as np
import pytest

from conftest import REPO, golden


def test_library_exports_every_declared_symbol():
    from numbskull_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(REPO, "include", "numbskull_b200.h")).read()
    declared = set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.nb_abi_version() == 1


def test_record_layouts_match_header():
    from numbskull_b200 import numbskulltypes as t
    header = open(os.path.join(REPO, "include", "numbskull_b200.h")).read()
    for name, size in (("nb_weight_rec", 9), ("nb_variable_rec", 27), ("nb_factor_rec", 34),
                       ("nb_ftv_rec", 16), ("nb_vtf_rec", 24)):
        assert re.search(r"%s;\s*/\*\s*%d B" % (name, size), header), name
    assert t.Variable.names == ("isEvidence", "initialValue", "dataType", "cardinality", "vtf_offset")
    assert t.Factor.names == ("factorFunction", "weightId", "featureValue", "arity", "ftv_offset")


def test_loaders_match_reference_on_coin_graph(tmp_path):
    import numbskull_b200 as nb
    z = golden("coin")
    for n in ("meta", "weights", "variables", "factors"):
        z["raw_" + n].tofile(str(tmp_path / ("graph." + n)))
    ns = nb.numbskull.load([str(tmp_path), "-q"])
    fg = ns.factorGraphs[0]
    for k in ("weight", "variable", "factor", "fmap", "vmap", "factor_index"):
        assert np.array_equal(getattr(fg, k), z[k]), k
    assert np.array_equal(fg.cstart, np.arange(19))
    assert fg.count.shape == (18,) and fg.var_value.shape == (1, 18)


@pytest.mark.parametrize("name", ["bool", "cat", "lf", "ising"])
def test_compute_var_map_matches_reference(name):
    import numbskull_b200 as nb
    z = golden("varmap_" + name)
    ns = nb.NumbSkull(quiet=True)
    v = z["variable"].copy()
    v["vtf_offset"] = -1
    ns.loadFactorGraph(z["weight"].copy(), v, z["factor"].copy(), z["fmap"].copy(),
                       z["domain_mask"].copy(), len(z["fmap"]))
    fg = ns.factorGraphs[0]
    assert np.array_equal(fg.variable, z["variable"])
    assert np.array_equal(fg.vmap, z["vmap"])
    for b in z["vmap"]:
        s, n = b["factor_index_offset"], b["factor_index_length"]
        assert np.array_equal(fg.factor_index[s:s + n], z["factor_index"][s:s + n])


def test_compute_var_map_matches_oracle_on_random_graphs(oracle):
    from numbskull_b200 import synth
    from numbskull_b200.dataloading import assign_vtf_offsets, compute_var_map
    from numbskull_b200.numbskulltypes import VarToFactor
    rng = np.random.default_rng(31)
    for trial in range(20):
        g = synth.random_graph(int(rng.integers(1, 60)), int(rng.integers(0, 150)), rng,
                               funcs=(0, 1, 2, 3, 12, 14), card=int(rng.integers(2, 6)),
                               categorical_frac=float(rng.random()), allow_repeats=True)
        w, v, f, fm, dm, e = g
        v1, v2 = v.copy(), v.copy()
        n = assign_vtf_offsets(v1)
        v2["vtf_offset"] = np.concatenate(([0], np.cumsum(np.where(v["dataType"] == 0, 1, v["cardinality"]))[:-1]))
        assert np.array_equal(v1, v2)
        vm1, vm2 = np.zeros(n, VarToFactor), np.zeros(n, VarToFactor)
        fi1, fi2 = np.zeros(len(fm), np.int64), np.zeros(len(fm), np.int64)
        compute_var_map(v1, f, fm, vm1, fi1, dm)
        oracle.compute_var_map(v2, f, fm, vm2, fi2, dm)
        assert np.array_equal(vm1, vm2)
        for b in vm1:
            s, k = b["factor_index_offset"], b["factor_index_length"]
            assert np.array_equal(fi1[s:s + k], fi2[s:s + k])


def test_compute_var_map_threaded_path_matches_oracle(oracle):
    """> 65536 fmap entries: the multi-threaded scatter + per-bucket sort must give the same
    (ascending, de-duplicated) buckets as the sequential reference algorithm."""
    from numbskull_b200 import synth
    from numbskull_b200.dataloading import assign_vtf_offsets, compute_var_map
    from numbskull_b200.numbskulltypes import VarToFactor
    w, v, f, fm, dm, e = synth.random_graph(20000, 90000, np.random.default_rng(77), funcs=(0, 1, 2, 3, 12, 14),
                                            card=5, categorical_frac=0.4, allow_repeats=True, max_arity=4)
    assert len(fm) > (1 << 16)
    v1, v2 = v.copy(), v.copy()
    n = assign_vtf_offsets(v1)
    assign_vtf_offsets(v2)
    vm1, vm2 = np.zeros(n, VarToFactor), np.zeros(n, VarToFactor)
    fi1, fi2 = np.zeros(len(fm), np.int64), np.zeros(len(fm), np.int64)
    compute_var_map(v1, f, fm, vm1, fi1, dm)
    oracle.compute_var_map(v2, f, fm, vm2, fi2, dm)
    assert np.array_equal(vm1, vm2)
    keep = np.repeat(np.arange(len(vm1)), vm1["factor_index_length"])
    pos = np.repeat(vm1["factor_index_offset"], vm1["factor_index_length"]) + \
        (np.arange(len(keep)) - np.repeat(np.cumsum(vm1["factor_index_length"]) - vm1["factor_index_length"],
                                          vm1["factor_index_length"]))
    assert np.array_equal(fi1[pos], fi2[pos])


@pytest.mark.parametrize("kind", ["kbc", "categorical", "repeats"])
def test_compute_var_map_partitioned_path_is_bit_exact(oracle, kind, monkeypatch):
    """>= 2^20 fmap entries with factors tiling fmap in order: the atomics-free partitioned build
    (nb_host.cpp var_map_partitioned) must equal the general path byte for byte, gaps included,
    and the sequential oracle on every bucket."""
    from numbskull_b200 import synth
    from numbskull_b200.dataloading import assign_vtf_offsets, compute_var_map
    from numbskull_b200.numbskulltypes import VarToFactor
    rng = np.random.default_rng(5)
    if kind == "kbc":
        w, v, f, fm, dm, e = synth.kbc(230000, rng=rng, n_weights=1000)
    elif kind == "categorical":
        w, v, f, fm, dm, e = synth.categorical(180000, card=7, rng=rng, n_weights=1000)
    else:   # the same variable several times in one factor: duplicates inside a bucket
        w, v, f, fm, dm, e = synth.kbc(230000, rng=rng, n_weights=1000, window=2)
    assert len(fm) >= (1 << 20)
    v = v.copy()
    n = assign_vtf_offsets(v)
    out = {}
    for mode in ("partitioned", "3 threads", "16 threads", "61 threads", "general"):
        if mode == "general":
            monkeypatch.setenv("NUMBSKULL_B200_VARMAP_GENERAL", "1")
        elif mode != "partitioned":
            monkeypatch.setenv("NUMBSKULL_B200_HOST_THREADS", mode.split()[0])     # the range count follows it
        vm, fi = np.zeros(n, VarToFactor), np.zeros(len(fm), np.int64)
        compute_var_map(v, f, fm, vm, fi, dm)
        out[mode] = (vm, fi)
    for mode in out:
        assert np.array_equal(out[mode][0], out["general"][0]), mode
        assert np.array_equal(out[mode][1], out["general"][1]), mode
    vm2, fi2 = np.zeros(n, VarToFactor), np.zeros(len(fm), np.int64)
    oracle.compute_var_map(v.copy(), f, fm, vm2, fi2, dm)
    vm1, fi1 = out["partitioned"]
    assert np.array_equal(vm1, vm2)
    lens = vm1["factor_index_length"]
    pos = np.repeat(vm1["factor_index_offset"], lens) + (np.arange(lens.sum()) - np.repeat(np.cumsum(lens) - lens, lens))
    assert np.array_equal(fi1[pos], fi2[pos])
    if kind == "repeats":
        assert (vm1["factor_index_length"].sum() < len(fm))       # duplicates were dropped


def test_compute_var_map_partitioned_path_with_fewer_variables_than_threads(monkeypatch):
    """70 variables, 1.1 M fmap entries: most ranges are empty, every bucket is a hub."""
    from numbskull_b200.dataloading import assign_vtf_offsets, compute_var_map
    from numbskull_b200.numbskulltypes import Variable, Factor, FactorToVar, VarToFactor
    rng = np.random.default_rng(0)
    nv, nf = 70, 550000
    v = np.zeros(nv, Variable)
    v["cardinality"] = rng.integers(2, 5, nv)
    v["dataType"] = rng.integers(0, 2, nv)
    f = np.zeros(nf, Factor)
    f["arity"], f["ftv_offset"], f["factorFunction"] = 2, np.arange(nf) * 2, 12
    fm = np.zeros(2 * nf, FactorToVar)
    fm["vid"] = rng.integers(0, nv, 2 * nf)
    fm["dense_equal_to"] = rng.integers(0, 2, 2 * nf)
    n = assign_vtf_offsets(v)
    res = {}
    for mode in ("default", "33", "general"):
        if mode == "general":
            monkeypatch.setenv("NUMBSKULL_B200_VARMAP_GENERAL", "1")
        elif mode != "default":
            monkeypatch.setenv("NUMBSKULL_B200_HOST_THREADS", mode)
        vm, fi = np.zeros(n, VarToFactor), np.zeros(len(fm), np.int64)
        compute_var_map(v, f, fm, vm, fi, np.zeros(nv, bool))
        res[mode] = (vm, fi)
    for mode in res:
        assert np.array_equal(res[mode][0], res["general"][0]) and np.array_equal(res[mode][1], res["general"][1]), mode


def test_compute_var_map_partitioned_path_reports_bad_graphs():
    from numbskull_b200 import synth
    from numbskull_b200.dataloading import assign_vtf_offsets, compute_var_map
    from numbskull_b200.numbskulltypes import VarToFactor
    w, v, f, fm, dm, e = synth.kbc(230000, rng=np.random.default_rng(6), n_weights=1000)
    v = v.copy()
    n = assign_vtf_offsets(v)
    fm = fm.copy()
    fm["vid"][len(fm) // 2] = len(v) + 5
    with pytest.raises(Exception, match="vid out of range"):
        compute_var_map(v, f, fm, np.zeros(n, VarToFactor), np.zeros(len(fm), np.int64), dm)


def test_factors_to_skip_is_bounds_checked():
    """The reference overruns factor_index here; we size it safely and skip correctly."""
    import numbskull_b200 as nb
    from numbskull_b200 import synth
    w, v, f, fm, dm, e = synth.random_graph(25, 60, np.random.default_rng(4), allow_repeats=True)
    skip = np.array([0, 3, 4, 17, 59], np.int64)
    ns = nb.NumbSkull(quiet=True)
    ns.loadFactorGraph(w, v, f, fm, dm, e, factors_to_skip=skip)
    fg = ns.factorGraphs[0]
    used = set()
    for b in fg.vmap:
        s, n = b["factor_index_offset"], b["factor_index_length"]
        used |= set(fg.factor_index[s:s + n].tolist())
    assert not (used & set(skip.tolist()))
    assert used == set(range(60)) - set(skip.tolist()) - {i for i in range(60) if f["arity"][i] == 0}


def test_deepdive_writer_round_trip(tmp_path):
    """synth.write_deepdive -> loadFGFromFile gives back the same records
    (incl. an explicit graph.domains translation, dataloading.py:159-237)."""
    import numbskull_b200 as nb
    from numbskull_b200 import synth
    w, v, f, fm, dm, e = synth.random_graph(30, 70, np.random.default_rng(12), funcs=(0, 1, 3, 12, 15),
                                            card=4, categorical_frac=0.5, feature_values=True)
    doms = {int(i): np.array([3, 10, 11, 40]) for i in np.nonzero(v["dataType"] == 1)[0][:4]}
    v_file = v.copy()
    for i, d in doms.items():
        v_file["initialValue"][i] = d[v["initialValue"][i]]
    synth.write_deepdive(str(tmp_path), w, v_file, f, fm, domains=doms)
    ns = nb.numbskull.load([str(tmp_path), "-q"])
    fg = ns.factorGraphs[0]
    assert np.array_equal(fg.weight, w)
    for k in ("isEvidence", "initialValue", "dataType", "cardinality"):
        assert np.array_equal(fg.variable[k], v[k]), k
    for k in ("factorFunction", "weightId", "featureValue", "arity", "ftv_offset"):
        assert np.array_equal(fg.factor[k], f[k]), k
    assert np.array_equal(fg.fmap, fm)
    for i, d in doms.items():
        s = fg.variable["vtf_offset"][i]
        assert np.array_equal(fg.vmap["value"][s:s + 4], d)


def test_option_table_matches_reference_defaults():
    import numbskull_b200 as nb
    ns = nb.NumbSkull()
    want = dict(n_learning_epoch=0, n_inference_epoch=0, stepsize=0.01, decay=0.95, reg_param=0.01,
                regularization=2, truncation=1, burn_in=0, nthreads=1, sample_evidence=True,
                learn_non_evidence=False, quiet=False, verbose=False, directory='.', output_dir='.',
                metafile='graph.meta', weightfile='graph.weights', variablefile='graph.variables',
                factorfile='graph.factors', domainfile='graph.domains', dburl='')
    for k, v in want.items():
        assert getattr(ns, k) == v, k
    ns = nb.NumbSkull(n_learning_epoch=7, bogus=1)
    assert ns.n_learning_epoch == 7 and not hasattr(ns, "bogus")
    assert nb.inference.FACTORS["DP_GEN_LF_ACCURACY"] == 21 and nb.inference.FUNC_EQUAL == 3
    assert len(nb.inference.FACTORS) == 25


def test_import_numbskull_alias():
    """`import numbskull` (the reference's package name) resolves to this implementation, with the
    import forms the reference's own scripts use (test.py:5, loadfg.py:6-7, test_lf_learning.py:6-7)."""
    import numbskull
    from numbskull import numbskull as mod
    from numbskull.numbskulltypes import Weight, Variable, Factor, FactorToVar  # noqa: F401
    import numbskull_b200
    assert numbskull.NumbSkull is numbskull_b200.NumbSkull is mod.NumbSkull
    assert numbskull.inference.FACTORS["IMPLY_NATURAL"] == 0 and hasattr(mod, "load") and hasattr(mod, "main")
    assert Weight.itemsize == 9


def test_dump_formats(tmp_path):
    from numbskull_b200.factorgraph import FactorGraph
    z = golden("run_cat")
    fg = FactorGraph(z["weight"].copy(), z["variable"].copy(), z["factor"].copy(), z["fmap"].copy(),
                     z["vmap"].copy(), z["factor_index"].copy(), 1, 1, 0, 1)
    fg.count[:] = z["count"]
    p = str(tmp_path / "m.txt")
    fg.dump_probabilities(p, 20)
    want = []
    for i, v in enumerate(fg.variable):              # factorgraph.py:216-229 restated
        if v["cardinality"] == 2:
            want.append('%d %d %.3f\n' % (i, 1, float(fg.count[fg.cstart[i]]) / 20))
        else:
            for k in range(v["cardinality"]):
                want.append('%d %d %.3f\n' % (i, fg.vmap[v["vtf_offset"] + k]["value"],
                                              float(fg.count[fg.cstart[i] + k]) / 20))
    assert open(p).read() == ''.join(want)
    fg.dump_weights(p)
    assert open(p).read() == ''.join('%d %f\n' % (i, x) for i, x in enumerate(fg.weight_value[0]))


def test_no_cpu_fallback():
    """Without a CUDA device the hot path must fail loudly, never fall back."""
    from numbskull_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("CUDA device present")
    from numbskull_b200.factorgraph import FactorGraph
    z = golden("run_ising")
    fg = FactorGraph(z["weight"].copy(), z["variable"].copy(), z["factor"].copy(), z["fmap"].copy(),
                     z["vmap"].copy(), z["factor_index"].copy(), 1, 1, 0, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fg.inference(0, 1)


def test_product_never_imports_oracle():
    """Nothing under numbskull_b200/ may import, link or execute the oracle."""
    pkg = os.path.join(REPO, "numbskull_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.+\s*oracle|oracle[./\\]|libnb_oracle|\bnbo_")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(root, f)).read()
                assert not bad.search(src), os.path.join(root, f)


# --------------------------------------------------------------------------- edge cases of the host side
def test_empty_graph_host_side():
    """No variables / no factors: offsets, index build and the loaders are no-ops, not crashes."""
    from numbskull_b200 import dataloading
    from numbskull_b200.numbskulltypes import Weight, Variable, Factor, FactorToVar, VarToFactor
    variable, factor, fmap = np.zeros(0, Variable), np.zeros(0, Factor), np.zeros(0, FactorToVar)
    assert dataloading.assign_vtf_offsets(variable) == 0
    vmap, findex = np.zeros(0, VarToFactor), np.zeros(0, np.int64)
    dataloading.compute_var_map(variable, factor, fmap, vmap, findex, np.zeros(0, bool))
    empty = np.zeros(0, np.uint8)
    dataloading.load_weights(empty, 0, np.zeros(0, Weight))
    dataloading.load_variables(empty, 0, variable)
    dataloading.load_factors(empty, 0, factor, fmap, np.zeros(0, bool), variable, vmap)


def test_isolated_variables_and_zero_arity_factors(oracle):
    """Variables no factor touches get empty buckets; a factor without members contributes nothing."""
    from numbskull_b200 import dataloading, synth
    from numbskull_b200.numbskulltypes import VarToFactor
    w, v, f, fm, dm, e = synth.random_graph(12, 9, np.random.default_rng(4), max_arity=3)
    f = f.copy()
    f["arity"][4] = 0                                  # its fmap slots are simply never referenced
    used = np.zeros(len(v), bool)
    for x in f:
        used[fm["vid"][x["ftv_offset"]:x["ftv_offset"] + x["arity"]]] = True
    n_vtf = dataloading.assign_vtf_offsets(v)
    vmap, findex = np.zeros(n_vtf, VarToFactor), np.zeros(len(fm), np.int64)
    dataloading.compute_var_map(v, f, fm, vmap, findex, dm)
    want_vmap, want_index = np.zeros(n_vtf, VarToFactor), np.zeros(len(fm), np.int64)
    oracle.compute_var_map(v.copy(), f, fm, want_vmap, want_index, dm)
    assert np.array_equal(vmap, want_vmap)
    for b in vmap:
        s0, k = b["factor_index_offset"], b["factor_index_length"]
        assert np.array_equal(findex[s0:s0 + k], want_index[s0:s0 + k])
    assert 4 not in findex[:int(vmap["factor_index_length"].sum())]
    for i in np.nonzero(~used)[0]:
        assert vmap["factor_index_length"][v["vtf_offset"][i]] == 0


@pytest.mark.parametrize("which", ["weights", "variables", "factors"])
def test_truncated_files_are_rejected(which):
    """A short buffer is an error (the reference would read past the end of the array)."""
    from numbskull_b200 import dataloading
    from numbskull_b200.numbskulltypes import Weight, Variable, Factor, FactorToVar, VarToFactor
    z = golden("coin")
    raw = {n: z["raw_" + n] for n in ("weights", "variables", "factors")}
    nw, nv, nf, ne = len(z["weight"]), len(z["variable"]), len(z["factor"]), len(z["fmap"])
    data = raw[which][:len(raw[which]) - 3]
    with pytest.raises(Exception):
        if which == "weights":
            dataloading.load_weights(data, nw, np.zeros(nw, Weight))
        elif which == "variables":
            dataloading.load_variables(data, nv, np.zeros(nv, Variable))
        else:
            variable = z["variable"].copy()
            vmap = np.zeros(len(z["vmap"]), VarToFactor)
            dataloading.load_factors(data, nf, np.zeros(nf, Factor), np.zeros(ne, FactorToVar),
                                     np.zeros(nv, bool), variable, vmap)
