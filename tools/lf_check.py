"""Debug/validation helper: LF-model learning, GPU vs the CPU oracle at a given size."""
import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numbskull_b200 as nb, oracle
from numbskull_b200 import synth

def run(copies, n_lf, bv, epochs=4, step=1e-4, reg=1, lne=True, fix_w0=False):
    g = synth.lf_model(copies, n_lf, np.random.default_rng(1003))
    g[0]['isFixed'][0] = fix_w0
    ns = nb.NumbSkull(quiet=True); ns.loadFactorGraph(*g); fg = ns.factorGraphs[0]; fg.seed = 5
    fg.batch_visits = bv
    og = oracle.OracleGraph(fg.weight.copy(), fg.variable.copy(), fg.factor.copy(), fg.fmap.copy(), fg.vmap.copy(), fg.factor_index.copy(), 1, 3)
    fg.learn(0, epochs, step, 1.0, reg, 0.01, 1, learn_non_evidence=lne)
    og.learn(0, epochs, step, 1.0, reg, 0.01, 1, learn_non_evidence=lne)
    per = 1 + n_lf
    for name, vf, ve in (("gpu", fg.var_value[0], fg.var_value_evid[0]), ("oracle", og.var_value, og.var_value_evid)):
        yf, ye = vf[::per], ve[::per]
        lf = vf.reshape(copies, per)[:, 1:]
        agree = (lf == yf[:, None]).mean(); abst = (lf == 2).mean()
        print("   %-6s y_free mean %.3f  y_evid mean %.3f  mismatch %.4f  LF_free agree %.3f abstain %.3f" % (
            name, yf.mean(), ye.mean(), (yf != ye).mean(), agree, abst))
    d = np.abs(fg.weight_value[0] - og.weight_value)
    print("fix_w0 %d copies %d n_lf %d bv %d reg %d lne %d | gpu %s | oracle %s | maxdiff %.4f" % (
        fix_w0, copies, n_lf, bv, reg, lne, np.round(fg.weight_value[0][:5], 3), np.round(og.weight_value[:5], 3), d.max()))
    sys.stdout.flush()

if __name__ == "__main__":
    for args in [(20000, 100, 0, 4), (100000, 100, 0, 4), (100000, 100, 0, 8), (100000, 100, 5000, 4)]:
        run(*args)
