"""Host / device coherence of FactorGraph (numbskull_b200/factorgraph.py).

The reference mutates numpy arrays in place (numbskull/factorgraph.py:135-141,156-163,196-202),
so state persists between calls and callers may poke ``var_value`` / ``weight_value`` /
``count`` between them (salt/src/numbskull_master.py:213-224).  Here the working copy lives in
HBM: untouched arrays are not copied at call boundaries, touched ones round-trip.  Both modes
must be indistinguishable from the reference's behaviour.
"""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def _fg(z, seed=5):
    from numbskull_b200.factorgraph import FactorGraph
    return FactorGraph(z["weight"].copy(), z["variable"].copy(), z["factor"].copy(), z["fmap"].copy(),
                       z["vmap"].copy(), z["factor_index"].copy(), 1, 1, 0, 1, device=0, seed=seed)


def test_lazy_and_eager_mirrors_give_the_same_chain():
    """Three inference calls: one graph never shows its arrays until the end (lazy), the other hands
    them out after every call (eager round trips).  Same seed => same samples, counts, marginals."""
    z = golden("run_bool_l2")
    lazy, eager = _fg(z), _fg(z)
    for call in range(3):
        lazy.inference(2, 40, sample_evidence=True)
        eager.inference(2, 40, sample_evidence=True)
        _ = eager.var_value, eager.count, eager.marginals        # exposes the arrays: eager from now on
        assert "var_value" in lazy._stale and "count" in lazy._stale
        assert "var_value" not in eager._stale or call == 0
    assert np.array_equal(lazy.var_value, eager.var_value)
    assert np.array_equal(lazy.count, eager.count)
    assert np.array_equal(lazy.marginals, eager.marginals)
    assert np.allclose(lazy.marginals, lazy.count / 40.0)        # cumulative count / epochs of the last call
    assert lazy.count.max() > 40                                 # ... and it IS cumulative


def test_marginals_do_not_need_count():
    z = golden("run_bool_l2")
    fg = _fg(z)
    fg.inference(1, 30, sample_evidence=True)
    m = fg.marginals.copy()
    assert "count" in fg._stale                                  # the int64 array was never materialised
    cc = fg.counts_compact()
    assert cc.dtype == np.uint8 and np.array_equal(cc, fg.count)
    assert np.allclose(m, fg.count / 30.0)
    fg.inference(0, 300, sample_evidence=True)
    cc = fg.counts_compact()
    assert cc.dtype == np.uint16 and np.array_equal(cc, fg.count)
    assert np.allclose(fg.marginals, fg.count / 300.0)


def test_caller_edits_between_calls_are_seen():
    """var_value of an evidence variable is state the sweeps read but (sample_evidence=False) never
    write: an edit through the public array must reach the device; so must count and weight edits."""
    z = golden("run_bool_l2")
    ev = np.nonzero(z["variable"]["isEvidence"] == 1)[0]
    assert len(ev)
    a, b = _fg(z), _fg(z)
    a.inference(0, 5)
    b.inference(0, 5)
    flipped = 1 - a.var_value[0][ev]
    a.var_value[0][ev] = flipped                                  # edit in place through the handed-out array
    vv = b.var_value.copy()
    vv[0][ev] = flipped
    b.var_value = vv                                              # ... or by assignment
    a.inference(0, 50)
    b.inference(0, 50)
    assert np.array_equal(a.var_value[0][ev], flipped) and np.array_equal(a.var_value, b.var_value)
    assert np.array_equal(a.count, b.count)
    c = _fg(z)
    c.inference(0, 5)
    c.inference(0, 50)                                            # no edit: a different chain
    assert not np.array_equal(c.count, a.count)
    # count edits: zeroing count between calls restarts the tallies (count is cumulative otherwise)
    a.count[:] = 0
    a.inference(0, 10)
    assert a.count.max() <= 10 and np.allclose(a.marginals, a.count / 10.0)
    # weight edits: compared with the last uploaded copy, uploaded when different
    w = a.weight_value[0].copy()
    a.weight_value[0][:] = 0.0
    a.inference(0, 1)
    e0 = a.potentials()
    assert not e0.any()
    a.weight_value[0][:] = w
    assert a.potentials().any()


def test_learn_refreshes_weights_and_both_chains():
    z = golden("run_bool_l2")
    a, b = _fg(z, seed=3), _fg(z, seed=3)
    for fg in (a, b):
        fg.learn(2, 5, 0.01, 0.95, 2, 0.01, 1.0)
    _ = b.var_value, b.var_value_evid                             # b goes eager
    for fg in (a, b):
        fg.learn(0, 5, 0.01, 0.95, 2, 0.01, 1.0)
    assert np.array_equal(a.weight_value, b.weight_value)
    assert np.array_equal(a.var_value, b.var_value) and np.array_equal(a.var_value_evid, b.var_value_evid)
    assert not np.array_equal(a.weight_value[0], z["weight"]["initialValue"])


def test_invalidate_rebuilds_from_edited_records():
    """isEvidence is read when the device graph is built; invalidate() makes an edit take effect."""
    z = golden("run_bool_l2")
    fg = _fg(z)
    fg.inference(0, 20)
    q = np.nonzero(fg.variable["isEvidence"] == 0)[0][:3]
    before = fg.var_value[0][q].copy()
    fg.variable["isEvidence"][q] = 1
    fg.invalidate()
    fg.inference(0, 200)
    assert np.array_equal(fg.var_value[0][q], before)             # now evidence: never resampled
    assert fg.count.max() <= 220                                  # counts carried over (cumulative)
    fg.clear()
    assert not fg.count.any()
