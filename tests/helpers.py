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


def assert_grads_close(got, want, rtol=1e-4, row_atol=0.0, atol=0.0, what=""):
    """|got - want| <= rtol * |want| + row_atol * max_k |want[..., k]| + atol, element-wise, with a report of the
    worst element on failure.  `row_atol` is for gradients whose components are sums of cancelling terms (an
    anchor's four regression gradients come out of one GIoU expression: a component a thousand times smaller than
    its neighbours carries their rounding error, measured <= 4e-6 of the row maximum on B200)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    allowed = rtol * np.abs(want) + atol
    if row_atol:
        allowed = allowed + row_atol * np.abs(want).max(axis=-1, keepdims=True)
    err = np.abs(got - want)
    bad = ~(err <= allowed)          # NaN counts as bad
    if bad.any():
        i = np.unravel_index(np.argmax(np.where(bad, err - allowed, -np.inf)), err.shape)
        raise AssertionError("%s: %d of %d elements differ; worst at %s: got %r want %r (allowed %.3g)"
                             % (what, int(bad.sum()), bad.size, i, got[i], want[i], allowed[i]))


def check_losses_and_grads_against_oracle(b, asg, got_labels, losses, cls, reg, iou, rtol=1e-4):
    """Losses and all three gradient families of a device step against the oracle's stage 5 run on the DEVICE's
    PAA labels (identical to the oracle's own unless a documented tie flipped a positive set, in which case the
    caller has already checked that the flipped GTs are exempt).  Never skipped."""
    from oracle import paa_oracle
    use_iou = iou is not None
    forced = paa_oracle.with_labels(asg, got_labels, b.box_regression, b.gt_boxes)
    ref_losses, ref = paa_oracle.losses_and_grads(b.box_cls, b.box_regression, b.iou_pred if use_iou else None,
                                                  forced)
    np.testing.assert_allclose([float(x.detach()) for x in losses], [float(x) for x in ref_losses], rtol=rtol)
    assert_grads_close(flat_levels([t.grad for t in cls]), flat_levels(ref.box_cls), rtol=rtol, atol=1e-9,
                       what="grad box_cls")
    assert_grads_close(flat_levels([t.grad for t in reg]), flat_levels(ref.box_regression), rtol=rtol,
                       row_atol=1e-5, what="grad box_regression")
    if use_iou:
        g, r = flat_levels([t.grad for t in iou]), flat_levels(ref.iou_pred)
        # sigmoid(x) - target cancels where the prediction is on target: error relative to the largest element
        assert_grads_close(g, r, rtol=rtol, atol=2e-6 * float(np.abs(r).max()), what="grad iou_pred")
    return ref_losses
