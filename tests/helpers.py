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
