"""Pin the CPU oracle (oracle/nb_oracle.c) against arrays produced by the real
numba reference (tests/golden/make_golden.py).  Integer work and seeded
trajectories must be bit-exact."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden, golden_opts

RUNS = ["bool_l2", "bool_l1", "cat", "lf", "ising", "pairs", "allfuncs"]


def _og(oracle, z, seed):
    return oracle.OracleGraph(z["weight"].copy(), z["variable"].copy(), z["factor"].copy(),
                              z["fmap"].copy(), z["vmap"].copy(), z["factor_index"].copy(),
                              nthreads=1, seed=seed)


def test_rng_stream_matches_numba(oracle):
    z = golden("rng_stream")
    L = oracle.lib()
    st = C.create_string_buffer(L.nbo_sizeof_rng())
    L.nbo_mt_seed(st, int(z["seed"]))
    got = np.array([L.nbo_mt_double(st) for _ in range(len(z["np_rand"]))])
    assert np.array_equal(got, z["np_rand"])
    # numba's random.* and np.random.* are separate generators with the same stream
    assert np.array_equal(got, z["py_random"])


def test_truth_tables(oracle):
    z = golden("truth_tables")
    from numbskull_b200.numbskulltypes import Weight, Variable, Factor, FactorToVar, VarToFactor
    n_checked = 0
    for key in z.files:
        func, arity, card = (int(x[1:]) for x in key.split("_"))
        w = np.zeros(1, Weight)
        v = np.zeros(arity, Variable)
        v["cardinality"] = card
        v["vtf_offset"] = np.arange(arity)
        f = np.zeros(1, Factor)
        f["factorFunction"], f["arity"], f["featureValue"] = func, arity, 1.0
        fm = np.zeros(arity, FactorToVar)
        fm["vid"] = np.arange(arity)
        fm["dense_equal_to"] = (np.arange(arity) + 1) % card
        vm = np.zeros(arity, VarToFactor)
        og = oracle.OracleGraph(w, v, f, fm, vm, np.zeros(arity, np.int64))
        for row in z[key]:
            state = row[:arity].astype(np.int64)
            og.var_value[:] = state
            assert og.eval_factor(0, -1, 0) == row[arity], (key, state)
            i = arity + 1
            for m in range(arity):
                for k in range(card):
                    assert og.eval_factor(0, m, k) == row[i], (key, state, m, k)
                    i += 1
                    n_checked += 1
    assert n_checked > 1000


@pytest.mark.parametrize("name", ["bool", "cat", "lf", "ising"])
def test_compute_var_map(oracle, name):
    from numbskull_b200.numbskulltypes import VarToFactor
    z = golden("varmap_" + name)
    variable = z["variable"].copy()
    vmap = np.zeros(len(z["vmap"]), VarToFactor)
    findex = np.zeros(len(z["factor_index"]), np.int64)
    oracle.compute_var_map(variable, z["factor"], z["fmap"], vmap, findex, z["domain_mask"])
    assert np.array_equal(vmap, z["vmap"])
    # entries past a bucket's de-duplicated length are scratch in the reference too
    for b in z["vmap"]:
        s, n = b["factor_index_offset"], b["factor_index_length"]
        assert np.array_equal(findex[s:s + n], z["factor_index"][s:s + n])


@pytest.mark.parametrize("name", RUNS)
def test_seeded_trajectory_bit_exact(oracle, name):
    z = golden("run_" + name)
    o = golden_opts(z)
    og = _og(oracle, z, int(z["seed"]))
    assert np.array_equal(og.potentials(), z["potentials_initial"])
    og.learn(o.get("burn_in", 0), o.get("n_learning_epoch", 0), o.get("stepsize", 0.01),
             o.get("decay", 0.95), o.get("regularization", 2), o.get("reg_param", 0.01),
             o.get("truncation", 1), o.get("learn_non_evidence", False))
    assert np.array_equal(og.weight_value, z["weight_after_learn"][0])
    assert np.array_equal(og.var_value, z["var_value_after_learn"][0])
    assert np.array_equal(og.var_value_evid, z["var_value_evid_after_learn"][0])
    og.inference(o.get("burn_in", 0), o.get("n_inference_epoch", 0), sample_evidence=True)
    assert np.array_equal(og.count, z["count"])
    assert np.array_equal(og.var_value, z["var_value"][0])
    assert np.array_equal(og.marginals, z["marginals"])


def test_coin_graph(oracle):
    z = golden("coin")
    og = _og(oracle, z, int(z["seed"]))
    og.learn(0, 10, 0.01, 0.95, 2, 0.01, 1, False)
    og.inference(0, 10, sample_evidence=True)
    assert np.array_equal(og.weight_value, z["weight_value"][0])
    assert np.array_equal(og.count, z["count"])
    assert np.array_equal(og.var_value, z["var_value"][0])


def test_exact_marginals_three_var_graph(oracle):
    """SURVEY.md section 8c pin: ISTRUE w=1 on v0; EQUAL(v0,v1) w=.5;
    IMPLY_NATURAL(v1,v2) w=.8; OR(v0,v2) w=-.3 -> (0.8716, 0.7541, 0.6248)."""
    from numbskull_b200.numbskulltypes import Weight, Variable, Factor, FactorToVar, VarToFactor
    w = np.zeros(4, Weight)
    w["initialValue"] = [1.0, 0.5, 0.8, -0.3]
    v = np.zeros(3, Variable)
    v["cardinality"] = 2
    v["vtf_offset"] = np.arange(3)
    f = np.zeros(4, Factor)
    f["factorFunction"] = [4, 3, 0, 1]
    f["weightId"] = np.arange(4)
    f["featureValue"] = 1
    f["arity"] = [1, 2, 2, 2]
    f["ftv_offset"] = [0, 1, 3, 5]
    fm = np.zeros(7, FactorToVar)
    fm["vid"] = [0, 0, 1, 1, 2, 0, 2]
    vm = np.zeros(3, VarToFactor)
    fi = np.zeros(7, np.int64)
    oracle.compute_var_map(v, f, fm, vm, fi, np.zeros(3, np.bool_))
    og = oracle.OracleGraph(w, v, f, fm, vm, fi)
    m = oracle.exact_marginals(og)
    assert np.allclose(m, [0.8716, 0.7541, 0.6248], atol=5e-5)
    og.inference(100, 20000, sample_evidence=True)
    assert np.abs(og.marginals - m).max() < 0.01


def test_coloring_policy_checker(oracle):
    """The colouring checker itself: valid colourings, checkerboard on a grid under the cap,
    hashed priorities above it."""
    from numbskull_b200 import synth
    _, v, f, fm, _, _ = synth.ising_grid(9, 7)
    col, mode = oracle.coloring.policy_coloring(v, f, fm, 0x5EED)
    rr, cc = np.divmod(np.arange(63), 7)
    assert mode == 1 and np.array_equal(col, (rr + cc) % 2)
    col, mode = oracle.coloring.policy_coloring(v, f, fm, 0x5EED, cap=14)
    assert mode == 0 and col.max() >= 2
    assert oracle.coloring.conflicts(v, f, fm, col) == 0
    assert np.array_equal(col, oracle.coloring.greedy_coloring(v, f, fm, 0x5EED))
