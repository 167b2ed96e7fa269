"""Helpers shared by the parity tests."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rel_l2(a, b):
    """||a - b||_2 / ||b||_2 : the relative gradient error used for end-to-end comparisons (the
    max-norm variant above is dominated by a handful of cancellation-prone entries of weight_v
    gradients, whose weight-norm backward subtracts two nearly equal vectors)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def max_abs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def check_grads_against_golden(gold, grads, rtol, prefix="grad:"):
    """``grads``: name -> tensor.  weight_v entries are compared through the compressed views
    stored by oracle/make_golden.py (first 4 rows, row sums, column sums)."""
    worst = {}
    for k, g in grads.items():
        if k.endswith("weight_v"):
            views = {":rows4": g[:4], ":rowsum": g.sum(1), ":colsum": g.sum(0)}
            for suf, v in views.items():
                ref = gold[prefix + k + suf]
                worst[k + suf] = rel_err(v, ref)
        else:
            key = prefix + k
            if key not in gold:
                continue
            worst[k] = rel_err(g, gold[key])
    bad = {k: v for k, v in worst.items() if not v <= rtol}
    assert not bad, "gradient mismatch vs golden: %s" % bad
    return worst
