#!/usr/bin/env python
"""NumbSkull API object and the ``numbskull`` command line (reference:
numbskull/numbskull.py).  Option names, defaults, load paths, output files and
printed messages are the reference's; the sampler underneath is the CUDA
library (see factorgraph.py)."""
from __future__ import print_function

import argparse
import os
import sys

import numpy as np

from .dataloading import (assign_vtf_offsets, compute_var_map, dataType, load_domains,
                          load_factors, load_variables, load_weights)
from .factorgraph import FactorGraph
from .numbskulltypes import (Factor, FactorToVar, Meta, VarToFactor, Variable, Weight)

long = int  # `from past.builtins import long` in the reference

# (flags, dest, metavar, default, type, help) -- numbskull.py:18-126
_OPTIONS = [
    (("-o", "--output_dir"), "output_dir", "OUTPUT_DIR", ".", str,
     "Output dir to contain inference_result.out.text and inference_result.out.weights.text"),
    (("-m", "--meta", "--fg_meta"), "metafile", "META_FILE", "graph.meta", str,
     "factor graph metadata file"),
    (("-w", "--weight", "--weights"), "weightfile", "WEIGHTS_FILE", "graph.weights", str,
     "factor weight file"),
    (("-v", "--variable", "--variables"), "variablefile", "VARIABLES_FILE", "graph.variables", str,
     "factor graph variables file"),
    (("-f", "--factor", "--factors"), "factorfile", "FACTORS_FILE", "graph.factors", str,
     "factor file"),
    (("--domain", "--domains"), "domainfile", "DOMAINS_FILE", "graph.domains", str, "domain file"),
    (("-l", "--n_learning_epoch"), "n_learning_epoch", "NUM_LEARNING_EPOCHS", 0, int,
     "number of learning epochs"),
    (("-i", "--n_inference_epoch"), "n_inference_epoch", "NUM_INFERENCE_EPOCHS", 0, int,
     "number of inference epochs"),
    (("-s", "--stepsize", "-a", "--alpha"), "stepsize", "LEARNING_STEPSIZE", 0.01, float,
     "stepsize for learning"),
    (("-d", "--decay", "--diminish"), "decay", "LEARNING_DECAY", 0.95, float,
     "decay for updating stepsize during learning"),
    (("-r", "--reg_param"), "reg_param", "LEARNING_REGULARIZATION_PARAM", 0.01, float,
     "regularization penalty"),
    (("--regularization",), "regularization", "REGULARIZATION", 2, int,
     'regularization (l1 or l2) [Enter as "1" or "2"]'),
    (("-k", "--truncation"), "truncation", "TRUNCATION", 1, int,
     "If using l1 regularization, truncation is applied with probability 1/k and with magnitude "
     "step_size * reg_param * k. If not using l1 regularization, this parameter has no effect."),
    (("-b", "--burn_in"), "burn_in", "BURN_IN", 0, int, "number of burn-in epochs"),
    (("-t", "--threads", "--n_threads"), "nthreads", "NUM_THREADS", 1, int,
     "number of threads to be used (accepted for compatibility; the GPU path ignores it)"),
    (("-u", "--dburl"), "dburl", "DATABASE_URL", "", str,
     "url to database holding factor graph information"),
]

arguments = [(("directory",), {"metavar": "DIRECTORY", "nargs": "?", "default": ".", "type": str,
                               "help": "specify the directory of factor graph files"})]
arguments += [(names, {"metavar": metavar, "dest": dest, "default": default, "type": typ, "help": hlp})
              for names, dest, metavar, default, typ, hlp in _OPTIONS]

# numbskull.py:128-149 (sample_evidence is store_true with default True, as in the reference)
flags = [
    (("--sample_evidence",), {"default": True, "dest": "sample_evidence", "action": "store_true",
                              "help": "sample evidence variables"}),
    (("--learn_non_evidence",), {"default": False, "dest": "learn_non_evidence", "action": "store_true",
                                 "help": "learn from non-evidence variables"}),
    (("-q", "--quiet"), {"default": False, "dest": "quiet", "action": "store_true", "help": "quiet"}),
    (("--verbose",), {"default": False, "dest": "verbose", "action": "store_true", "help": "verbose"}),
]


class NumbSkull(object):
    """Main class (numbskull.py:152-391): holds the run options and the list of
    factor graphs."""

    def __init__(self, **kwargs):
        for names, opts in arguments + flags:
            dest = opts.get("dest", names[0])
            setattr(self, dest, kwargs.get(dest, opts["default"]))
        self.factorGraphs = []

    def _add(self, weight, variable, factor, fmap, vmap, factor_index, var_copies, weight_copies):
        fg = FactorGraph(weight, variable, factor, fmap, vmap, factor_index, var_copies,
                         weight_copies, len(self.factorGraphs), self.nthreads)
        self.factorGraphs.append(fg)
        return fg

    def loadFactorGraphRaw(self, weight, variable, factor, fmap, vmap, factor_index,
                           var_copies=1, weight_copies=1):
        """numbskull.py:183-190."""
        self._add(weight, variable, factor, fmap, vmap, factor_index, var_copies, weight_copies)

    def loadFactorGraph(self, weight, variable, factor, fmap, domain_mask, edges,
                        var_copies=1, weight_copies=1, factors_to_skip=np.empty(0, np.int64)):
        """numbskull.py:192-243.  Note: factors_to_skip must be sorted."""
        assert(type(weight) == np.ndarray and weight.dtype == Weight)
        assert(type(variable) == np.ndarray and variable.dtype == Variable)
        assert(type(factor) == np.ndarray and factor.dtype == Factor)
        assert(type(fmap) == np.ndarray and fmap.dtype == FactorToVar)
        assert(type(domain_mask) == np.ndarray and domain_mask.dtype == np.bool_)
        assert(type(edges) == int or type(edges) == long or type(edges) == np.int64)
        assert(type(factors_to_skip) == np.ndarray and factors_to_skip.dtype == np.int64)

        # `edges` is recomputed exactly as the reference does (:217)
        edges = int(factor["arity"].sum()) - int(factor[factors_to_skip]["arity"].sum())
        num_vtfs = assign_vtf_offsets(variable)
        vmap = np.zeros(num_vtfs, VarToFactor)
        factor_index = np.zeros(edges, np.int64)
        compute_var_map(variable, factor, fmap, vmap, factor_index, domain_mask, factors_to_skip)
        self._add(weight, variable, factor, fmap, vmap, factor_index, var_copies, weight_copies)

    def loadFGFromFile(self, directory=None, metafile=None, weightfile=None, variablefile=None,
                       factorfile=None, domainfile=None, var_copies=1, weight_copies=1):
        """numbskull.py:245-353: DeepDive binary graph.{meta,weights,variables,domains,factors}."""
        if not self.directory:
            print("No factor graph specified")
            return
        directory = self.directory
        metafile = metafile or self.metafile
        weightfile = weightfile or self.weightfile
        variablefile = variablefile or self.variablefile
        factorfile = factorfile or self.factorfile
        domainfile = domainfile or self.domainfile
        print_info = not self.quiet
        print_only_meta = not self.verbose

        # graph.meta may carry extra columns (test/graph.meta has 8); the first 4 count
        meta = np.loadtxt(os.path.join(directory, metafile), delimiter=',', dtype=Meta,
                          usecols=(0, 1, 2, 3))
        meta = meta[()]
        if print_info:
            print("Meta:")
            print("    weights:  ", meta["weights"])
            print("    variables:", meta["variables"])
            print("    factors:  ", meta["factors"])
            print("    edges:    ", meta["edges"])
            print()

        weight = np.zeros(meta["weights"], Weight)
        load_weights(np.fromfile(os.path.join(directory, weightfile), np.uint8), meta["weights"], weight)
        if print_info and not print_only_meta:
            print("Weights:")
            for (i, w) in enumerate(weight):
                print("    weightId:", i)
                print("        isFixed:", w["isFixed"])
                print("        weight: ", w["initialValue"])
            print()

        variable = np.zeros(meta["variables"], Variable)
        load_variables(np.fromfile(os.path.join(directory, variablefile), np.uint8),
                       meta["variables"], variable)
        sys.stdout.flush()
        if print_info and not print_only_meta:
            print("Variables:")
            for (i, v) in enumerate(variable):
                print("    variableId:", i)
                print("        isEvidence:  ", v["isEvidence"])
                print("        initialValue:", v["initialValue"])
                print("        dataType:    ", v["dataType"], "(", dataType(v["dataType"]), ")")
                print("        cardinality: ", v["cardinality"])
                print()

        num_vtfs = assign_vtf_offsets(variable)
        print("#VTF = %s" % num_vtfs)
        sys.stdout.flush()

        vmap = np.zeros(num_vtfs, VarToFactor)
        factor_index = np.zeros(meta["edges"], np.int64)

        domain_mask = np.zeros(meta["variables"], np.bool_)
        domain_file = os.path.join(directory, domainfile)
        if os.path.isfile(domain_file) and os.stat(domain_file).st_size > 0:
            load_domains(np.fromfile(domain_file, np.uint8), domain_mask, vmap, variable)
            sys.stdout.flush()

        factor = np.zeros(meta["factors"], Factor)
        fmap = np.zeros(meta["edges"], FactorToVar)
        load_factors(np.fromfile(os.path.join(directory, factorfile), np.uint8), meta["factors"],
                     factor, fmap, domain_mask, variable, vmap)
        sys.stdout.flush()

        compute_var_map(variable, factor, fmap, vmap, factor_index, domain_mask)
        print("COMPLETED VMAP INDEXING")
        sys.stdout.flush()

        self._add(weight, variable, factor, fmap, vmap, factor_index, var_copies, weight_copies)

    def getFactorGraph(self, fgID=0):
        return self.factorGraphs[fgID]

    def inference(self, fgID=0, out=True):
        """numbskull.py:359-371."""
        fg = self.factorGraphs[fgID]
        fg.inference(self.burn_in, self.n_inference_epoch, sample_evidence=self.sample_evidence,
                     diagnostics=not self.quiet)
        if out:
            fg.dump_probabilities(os.path.join(self.output_dir, "inference_result.out.text"),
                                  self.n_inference_epoch)

    def learning(self, fgID=0, out=True):
        """numbskull.py:373-391."""
        fg = self.factorGraphs[fgID]
        fg.learn(self.burn_in, self.n_learning_epoch, self.stepsize, self.decay, self.regularization,
                 self.reg_param, self.truncation, diagnostics=not self.quiet, verbose=self.verbose,
                 learn_non_evidence=self.learn_non_evidence)
        if out:
            fg.dump_weights(os.path.join(self.output_dir, "inference_result.out.weights.text"))


def load(argv=None):
    """numbskull.py:394-416."""
    if argv is None:
        argv = sys.argv[1:]
    parser = argparse.ArgumentParser(description="Runs a Gibbs sampler", epilog="")
    parser.add_argument("--version", action='version', version="%(prog)s 0.0",
                        help="print version number")
    for names, opts in arguments + flags:
        parser.add_argument(*names, **opts)
    args = parser.parse_args(argv)
    ns = NumbSkull(**vars(args))
    ns.loadFGFromFile()
    return ns


def main(argv=None):
    """numbskull.py:419-423."""
    ns = load(argv)
    ns.learning()
    ns.inference()
