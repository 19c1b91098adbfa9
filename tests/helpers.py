"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np
import torch

from oracle import make_golden
from paa_b200 import synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def loss_case_batch(name):
    kw = dict(make_golden.LOSS_CASES)[name]
    return synthetic.make_batch(**kw)


def post_case_batch(name):
    kw = dict(make_golden.POST_CASES)[name]
    return synthetic.make_inference_batch(**kw)


def flat_levels(ts):
    """list of [N,C,H,W] -> numpy [N, A, C] in the reference's flattened anchor order."""
    return np.concatenate([t.detach().cpu().permute(0, 2, 3, 1).reshape(t.shape[0], -1, t.shape[1]).numpy()
                           for t in ts], axis=1)


from paa_b200.synthetic import to_device_inputs  # noqa: E402,F401  (re-exported for the tests)


def topk_tie_exempt(oracle_asg, rel=1e-5):
    """Exemption (i) of SURVEY.md 8c: (image, gt) pairs for which, on some level, the k-th and
    (k+1)-th smallest oracle loss among the GT's anchors are closer than `rel` relative, so the
    top-k membership is decided by rounding noise."""
    exempt = set()
    prm = oracle_asg.params
    N, A = oracle_asg.N, oracle_asg.A
    loss = oracle_asg.combined_loss.numpy()
    matched = oracle_asg.matched_idx.numpy()
    for i in range(N):
        n_gt = int(matched[i].max()) + 1
        start = 0
        for n_l in oracle_asg.level_sizes:
            m = matched[i, start:start + n_l]
            v = loss[i, start:start + n_l]
            for g in range(n_gt):
                x = np.sort(v[m == g])
                if x.shape[0] > prm.topk:
                    a, b = x[prm.topk - 1], x[prm.topk]
                    if abs(b - a) <= rel * max(abs(a), abs(b)):
                        exempt.add((i, g))
            start += n_l
    return exempt


def gmm_tie_exempt(oracle_asg, abs_tol=1e-5):
    """Exemption (ii): GTs whose two best foreground scores are within abs_tol (structural ties), and
    (ii-b) GTs with a sample whose two weighted component log-probabilities are within abs_tol (the
    foreground / background call of `predict` is rounding noise, e.g. two coinciding components)."""
    from oracle import gmm_oracle
    exempt = set()
    for i, recs in enumerate(oracle_asg.gmm_records):
        for r in recs:
            if r.get("fit") is None:
                continue
            if (gmm_oracle.structural_tie_margin(r["fit"]) < abs_tol
                    or gmm_oracle.component_margin(r["fit"], r["sorted_loss"]) < abs_tol):
                exempt.add((i, r["gt"]))
    return exempt
