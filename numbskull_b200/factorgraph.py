"""FactorGraph: the drop-in seam (reference: numbskull/factorgraph.py:27-229).

Same constructor signature, public arrays and methods as the reference class.
The epoch loops that used to hand the arrays to the numba ``gibbsthread`` /
``learnthread`` through ``run_pool`` (factorgraph.py:13-24, :135-141, :156-163,
:196-202) call the CUDA library instead.  The numpy arrays stay the
caller-visible state: they are uploaded on entry to ``burnIn`` / ``inference``
/ ``learn`` and refreshed on exit, so code that pokes ``var_value`` or
``weight_value`` between calls (salt/src/numbskull_master.py:213-224) keeps
working.  ``workers`` (nthreads) is accepted and ignored on the GPU.
"""
from __future__ import print_function

import ctypes as C
import os
import sys

import numpy as np

from . import _lib
from .timer import Timer


def _default_seed():
    env = os.environ.get("NUMBSKULL_B200_SEED")
    if env is not None:
        return int(env) & 0xFFFFFFFFFFFFFFFF
    return int.from_bytes(os.urandom(8), "little")


class FactorGraph(object):
    """Device-backed factor graph with the reference's interface."""

    def __init__(self, weight, variable, factor, fmap, vmap, factor_index,
                 var_copies, weight_copies, fid, workers, device=None, seed=None):
        self.weight = weight
        self.variable = variable
        self.factor = factor
        self.fmap = fmap
        self.vmap = vmap
        self.factor_index = factor_index

        nvar = self.variable.shape[0]
        # exclusive cumsum of (1 if cardinality == 2 else cardinality), factorgraph.py:40-45
        per_var = self.variable["cardinality"].astype(np.int64)
        per_var[per_var == 2] = 1
        self.cstart = np.zeros(nvar + 1, np.int64)
        np.cumsum(per_var, out=self.cstart[1:])
        self.count = np.zeros(self.cstart[nvar], np.int64)

        init = self.variable["initialValue"].astype(np.int64)
        self.var_value_evid = np.tile(init, (var_copies, 1))
        self.var_value = np.tile(init, (var_copies, 1))
        self.weight_value = np.tile(self.weight["initialValue"].astype(np.float64),
                                    (weight_copies, 1))

        # scratch the reference exposes; kept for shape compatibility only
        maxcard = int(self.variable["cardinality"].max()) if nvar else 0
        self.Z = np.zeros((workers, maxcard))
        maxlen = int(self.vmap["factor_index_length"].max()) if self.vmap.size else 0
        self.fids = np.zeros((workers, 2 * maxlen), factor_index.dtype)

        self.fid = fid
        assert(workers > 0)
        self.threads = workers
        self.marginals = np.zeros(self.cstart[nvar])
        self.inference_epoch_time = 0.0
        self.inference_total_time = 0.0
        self.learning_epoch_time = 0.0
        self.learning_total_time = 0.0

        # ---- device side ----
        self.device = device if device is not None else int(os.environ.get("LOCAL_RANK", 0))
        self.seed = _default_seed() if seed is None else int(seed)
        self.color_seed = int(os.environ.get("NUMBSKULL_B200_COLOR_SEED", 0x5EED))
        self.batch_visits = 0          # learning mini-batch policy, 0 = library default
        self.global_vid = None         # set by the partitioner for multi-GPU graphs
        self.preset_color = None
        self.warp_row_words = int(os.environ.get("NUMBSKULL_B200_WARP_ROW_WORDS", 0))  # 0 = default
        self.sigma_shift = int(os.environ.get("NUMBSKULL_B200_SIGMA_SHIFT", 0))        # 0 = default
        self.deferred_coloring = False  # partitioned graphs: colouring is driven by partition.py
        self._g = None

    # ------------------------------------------------------------------
    # device graph management
    # ------------------------------------------------------------------
    def _device_graph(self):
        """Build (once) the HBM-resident form of the graph."""
        if self._g is None:
            L = _lib.lib()
            self._keep = dict(
                weight=np.ascontiguousarray(self.weight), variable=np.ascontiguousarray(self.variable),
                factor=np.ascontiguousarray(self.factor), fmap=np.ascontiguousarray(self.fmap),
                vmap=np.ascontiguousarray(self.vmap),
                factor_index=_lib.contiguous(self.factor_index, np.int64),
                global_vid=None if self.global_vid is None else _lib.contiguous(self.global_vid, np.int64),
                preset_color=None if self.preset_color is None else _lib.contiguous(self.preset_color, np.int32))
            k = self._keep
            desc = _lib.GraphDesc(
                _lib.ptr(k["weight"]), len(k["weight"]), _lib.ptr(k["variable"]), len(k["variable"]),
                _lib.ptr(k["factor"]), len(k["factor"]), _lib.ptr(k["fmap"]), len(k["fmap"]),
                _lib.ptr(k["vmap"]), len(k["vmap"]), _lib.ptr(k["factor_index"]), len(k["factor_index"]),
                self.device, self.color_seed, _lib.ptr(k["global_vid"]),
                int(self.warp_row_words), int(self.sigma_shift), _lib.ptr(k["preset_color"]),
                int(bool(self.deferred_coloring)))
            g = C.c_void_p()
            _lib.check(L.nb_graph_create(C.byref(desc), C.byref(g)))
            self._g = g
            self._keep = None
        return self._g

    def device_info(self):
        info = _lib.GraphInfo()
        _lib.check(_lib.lib().nb_graph_get_info(self._device_graph(), C.byref(info)))
        return info.as_dict()

    def colors(self):
        """Colour of every variable (original order); -1 for non-owned variables."""
        out = np.empty(self.variable.shape[0], np.int32)
        _lib.check(_lib.lib().nb_graph_get_colors(self._device_graph(), _lib.ptr(out)))
        return out

    def color_edges(self):
        n = self.device_info()["n_colors"]
        out = np.zeros(max(n, 1), np.int64)
        _lib.check(_lib.lib().nb_graph_color_edges(self._device_graph(), _lib.ptr(out)))
        return out[:n]

    def potentials(self, var_ids=None, evid_chain=False, var_copy=0, weight_copy=0):
        """Conditional energies potential(v, k) (inference.py:55-71) for every
        value of the given variables, concatenated per variable."""
        g = self._device_graph()
        self._upload(var_copy, weight_copy)
        if var_ids is None:
            var_ids = np.arange(self.variable.shape[0], dtype=np.int64)
        var_ids = _lib.contiguous(var_ids, np.int64)
        cards = self.variable["cardinality"][var_ids].astype(np.int64)
        offs = np.zeros(len(var_ids), np.int64)
        if len(var_ids) > 1:
            np.cumsum(cards[:-1], out=offs[1:])
        out = np.zeros(int(cards.sum()), np.float64)
        _lib.check(_lib.lib().nb_potentials(g, int(bool(evid_chain)), _lib.ptr(var_ids), len(var_ids),
                                            _lib.ptr(offs), _lib.ptr(out), len(out)))
        return out

    def _row(self, arr, copy, dtype):
        """The caller-visible row `arr[copy]` as a C-contiguous array the library can read AND
        write in place (true for arrays built by __init__; anything else goes through a copy)."""
        row = arr[copy]
        if row.dtype == dtype and row.flags.c_contiguous and row.flags.writeable:
            return row, None
        tmp = np.ascontiguousarray(row, dtype=dtype)
        return tmp, row

    def _upload(self, var_copy, weight_copy, evid=True):
        L, g = _lib.lib(), self._device_graph()
        _lib.check(L.nb_set_var_values(g, 0, _lib.ptr(self._row(self.var_value, var_copy, np.int64)[0])))
        if evid:
            _lib.check(L.nb_set_var_values(g, 1, _lib.ptr(self._row(self.var_value_evid, var_copy, np.int64)[0])))
        _lib.check(L.nb_set_weights(g, _lib.ptr(self._row(self.weight_value, weight_copy, np.float64)[0])))

    def _download(self, var_copy, weight_copy, evid=False, weights=False, counts=False, epochs=0):
        L, g = _lib.lib(), self._g

        def fetch(arr, copy, dtype, call):
            buf, dst = self._row(arr, copy, dtype)
            _lib.check(call(_lib.ptr(buf)))
            if dst is not None:
                dst[:] = buf

        fetch(self.var_value, var_copy, np.int64, lambda p: L.nb_get_var_values(g, 0, p))
        if evid:
            fetch(self.var_value_evid, var_copy, np.int64, lambda p: L.nb_get_var_values(g, 1, p))
        if weights:
            fetch(self.weight_value, weight_copy, np.float64, lambda p: L.nb_get_weights(g, p))
        if counts:
            assert self.count.dtype == np.int64 and self.count.flags.c_contiguous
            if epochs:
                # count += tally and marginals = count / epochs (factorgraph.py:172-173) in one host pass
                m = self.marginals
                if not (isinstance(m, np.ndarray) and m.dtype == np.float64 and m.shape == self.count.shape
                        and m.flags.c_contiguous and m.flags.writeable):
                    m = np.empty(self.count.shape, np.float64)
                _lib.check(L.nb_get_counts_marginals(g, _lib.ptr(self.count), 1, _lib.ptr(m), float(epochs)))
                self.marginals = m
            else:
                _lib.check(L.nb_get_counts(g, _lib.ptr(self.count), 1))

    def clear(self):
        """factorgraph.py:75-78 (the thread pool's role is played by the device graph)."""
        self.count[:] = 0
        if self._g is not None:
            _lib.lib().nb_graph_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            if getattr(self, "_g", None) is not None:
                _lib.lib().nb_graph_destroy(self._g)
                self._g = None
        except Exception:
            pass

    #################
    #    GETTERS    #
    #################

    def getWeights(self, weight_copy=0):
        return self.weight_value[weight_copy][:]

    def getMarginals(self, varIds=None):
        if not varIds:
            return self.marginals
        else:
            return self.marginals[varIds]

    #####################
    #    DIAGNOSTICS    #
    #####################

    def diagnostics(self, epochs):
        """Marginal histogram, same text as factorgraph.py:99-113."""
        print('Inference took %.03f sec.' % self.inference_total_time)
        epochs = epochs or 1
        bins = 10
        assert (self.count >= 0).all() and (self.count <= epochs).all()
        hist = np.bincount(np.minimum(self.count * bins // epochs, bins - 1), minlength=bins)
        for i in range(bins):
            start = i / 10.0
            end = (i + 1) / 10.0
            print("Prob. " + str(start) + ".." + str(end) + ": \
                  " + str(hist[i]) + " variables")

    def diagnosticsLearning(self, weight_copy=0):
        print('Learning epoch took %.03f sec.' % self.learning_epoch_time)
        print("Weights:")
        for (i, w) in enumerate(self.weight):
            print("    weightId:", i)
            print("        isFixed:", w["isFixed"])
            print("        weight: ", self.weight_value[weight_copy][i])
            print()

    ################################
    #    INFERENCE AND LEARNING    #
    ################################

    def _sweeps(self, epochs, burnin, sample_evidence):
        _lib.check(_lib.lib().nb_gibbs_sweeps(self._g, int(epochs), int(bool(burnin)),
                                              int(bool(sample_evidence)), self.seed))

    def burnIn(self, epochs, sample_evidence, diagnostics=False,
               var_copy=0, weight_copy=0, _resident=False):
        """factorgraph.py:129-143."""
        if diagnostics:
            print("FACTOR " + str(self.fid) + ": STARTED BURN-IN...")
        if not _resident:
            self._upload(var_copy, weight_copy, evid=False)
        self._sweeps(epochs, True, sample_evidence)
        if not _resident:
            self._download(var_copy, weight_copy)
        if diagnostics:
            print("FACTOR " + str(self.fid) + ": DONE WITH BURN-IN")

    def inference(self, burnin_epochs, epochs, sample_evidence=False,
                  diagnostics=False, var_copy=0, weight_copy=0):
        """factorgraph.py:145-175."""
        L = _lib.lib()
        self._upload(var_copy, weight_copy, evid=False)     # the Gibbs sweep only touches chain 0
        if burnin_epochs > 0:
            self.burnIn(burnin_epochs, sample_evidence, diagnostics=diagnostics,
                        var_copy=var_copy, weight_copy=weight_copy, _resident=True)
        if diagnostics:
            print("FACTOR " + str(self.fid) + ": STARTED INFERENCE")
        _lib.check(L.nb_reset_counts(self._g))
        if diagnostics:
            for ep in range(epochs):
                with Timer() as timer:
                    self._sweeps(1, False, sample_evidence)
                self.inference_epoch_time = timer.interval
                self.inference_total_time += timer.interval
                print('Inference epoch #%d took %.03f sec.' % (ep, self.inference_epoch_time))
        elif epochs > 0:
            with Timer() as timer:
                self._sweeps(epochs, False, sample_evidence)
            self.inference_epoch_time = timer.interval / epochs
            self.inference_total_time += timer.interval
        self._download(var_copy, weight_copy, counts=True, epochs=epochs)
        if diagnostics:
            print("FACTOR " + str(self.fid) + ": DONE WITH INFERENCE")
        if diagnostics:
            self.diagnostics(epochs)

    def learn(self, burnin_epochs, epochs, stepsize, decay, regularization,
              reg_param, truncation, diagnostics=False, verbose=False,
              learn_non_evidence=False, var_copy=0, weight_copy=0):
        """factorgraph.py:177-208."""
        L = _lib.lib()
        self._upload(var_copy, weight_copy)
        if burnin_epochs > 0:
            self.burnIn(burnin_epochs, True, diagnostics=diagnostics,
                        var_copy=var_copy, weight_copy=weight_copy, _resident=True)
        if diagnostics:
            print("FACTOR " + str(self.fid) + ": STARTED LEARNING")

        def run(n, step):
            s = C.c_double(step)
            _lib.check(L.nb_learn_sweeps(self._g, int(n), C.byref(s), float(decay), int(regularization),
                                         float(reg_param), float(truncation),
                                         int(bool(learn_non_evidence)), self.seed, int(self.batch_visits)))
            return s.value

        if diagnostics:
            for ep in range(epochs):
                print("FACTOR " + str(self.fid) + ": EPOCH #" + str(ep))
                print("Current stepsize = " + str(stepsize))
                if verbose:
                    self._download(var_copy, weight_copy, weights=True)
                    self.diagnosticsLearning(weight_copy)
                sys.stdout.flush()
                with Timer() as timer:
                    stepsize = run(1, stepsize)
                self.learning_epoch_time = timer.interval
                self.learning_total_time += timer.interval
        elif epochs > 0:
            with Timer() as timer:
                stepsize = run(epochs, stepsize)
            self.learning_epoch_time = timer.interval / epochs
            self.learning_total_time += timer.interval
        self._download(var_copy, weight_copy, evid=True, weights=True)
        if diagnostics:
            print("FACTOR " + str(self.fid) + ": DONE WITH LEARNING")

    def dump_weights(self, fout, weight_copy=0):
        """Dump <wid, weight> text file in DW format (factorgraph.py:210-214)."""
        w = self.weight_value[weight_copy]
        with open(fout, 'w') as out:
            out.write(''.join('%d %f\n' % (i, x) for i, x in enumerate(w.tolist())))

    def dump_probabilities(self, fout, epochs):
        """Dump <vid, value, prob> text file in DW format (factorgraph.py:216-229)."""
        epochs = epochs or 1
        card = self.variable["cardinality"].astype(np.int64)
        nvar = len(card)
        binary = card == 2
        per_var = np.where(binary, 1, card)
        vid = np.repeat(np.arange(nvar, dtype=np.int64), per_var)
        k = np.arange(len(vid), dtype=np.int64) - np.repeat(self.cstart[:-1], per_var)
        value = np.ones(len(vid), np.int64)
        nb = ~np.repeat(binary, per_var)
        if nb.any():
            vtf = np.repeat(self.variable["vtf_offset"].astype(np.int64), per_var)
            value[nb] = self.vmap["value"][vtf[nb] + k[nb]]
        prob = self.count.astype(np.float64) / epochs
        with open(fout, 'w') as out:
            step = 1 << 20
            for s in range(0, len(vid), step):
                out.write(''.join('%d %d %.3f\n' % t for t in zip(vid[s:s + step].tolist(),
                                                                  value[s:s + step].tolist(),
                                                                  prob[s:s + step].tolist())))
