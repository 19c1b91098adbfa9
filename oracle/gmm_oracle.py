"""TEST INFRASTRUCTURE ONLY -- the per-GT two-component 1-D Gaussian-mixture fit of PAA.

The arithmetic is third-party: the reference calls ``sklearn.mixture.GaussianMixture`` at
paa_core/modeling/rpn/paa/loss.py:197-203 (ctor, fit, predict, score_samples).  scikit-learn is
*unpinned* by the reference (requirements.txt:10 lists ``sklearn``; setup.py:15-24 omits it); this
image has scikit-learn 1.9.0, whose source was read at
``sklearn/mixture/_base.py`` (fit_predict :202-318, _e_step :314-332, _estimate_log_prob_resp
:552-583), ``sklearn/mixture/_gaussian_mixture.py`` (_estimate_gaussian_parameters :282-321,
covariances 'full' :168-197, _compute_precision_cholesky :323-385, _estimate_log_gaussian_prob
:490-553, _initialize :836-882, _m_step :883-899) and ``sklearn/utils/_array_api.py`` (_logsumexp
:1338-1366).

Two implementations live here:

* ``impl="sklearn"``  -- run scikit-learn itself, with exactly the constructor arguments of
  loss.py:193-200.  This *is* the reference behaviour and is the default oracle.
* ``impl="numpy"``    -- a dtype-annotated restatement of what that call executes (float32 input,
  float64 responsibilities, float32 variance / precision after the first M-step, the tolerance
  stop on the mean log-likelihood, sklearn's max-masked log-sum-exp).  It documents the contract
  the CUDA kernel implements and is checked against ``impl="sklearn"`` in
  tests/test_oracle_gmm.py (identical iteration counts, components and positive counts).

Parity pin: the reference's tests hold no vector at this boundary -- the pin is
"scikit-learn 1.9.0 run in the same process", plus the recorded fits in tests/golden/.
"""
import math
import warnings

import numpy as np

MAX_ITER = 100       # sklearn default
TOL = 1e-3           # sklearn default
REG_COVAR = 1e-6     # sklearn default
_LOG_2PI = math.log(2.0 * math.pi)
_EPS10 = 10.0 * np.finfo(np.float64).eps


def _fit_sklearn(x):
    import sklearn.mixture as skm
    from sklearn.exceptions import ConvergenceWarning
    lo, hi = x.min(), x.max()                                   # loss.py:193
    gmm = skm.GaussianMixture(2, weights_init=[0.5, 0.5], means_init=[[lo], [hi]],
                              precisions_init=[[[1.0]], [[1.0]]])        # loss.py:194-200
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", ConvergenceWarning)
        gmm.fit(x)                                              # loss.py:201
    return dict(weights=np.asarray(gmm.weights_, np.float64).reshape(2),
                means=np.asarray(gmm.means_, np.float64).reshape(2),
                variances=np.asarray(gmm.covariances_, np.float64).reshape(2),
                n_iter=int(gmm.n_iter_), converged=bool(gmm.converged_),
                components=gmm.predict(x).astype(np.int64),     # loss.py:202
                scores=gmm.score_samples(x).astype(np.float64))  # loss.py:203


def _weighted_log_prob(x32, w, mu, pc):
    """a[i,k] = log N(x_i | mu_k, 1/pc_k^2) + log w_k with sklearn's dtype flow.
    ``pc`` is float64 before the first M-step and float32 afterwards; that decides where the
    products and the log-determinant are rounded (_estimate_log_gaussian_prob)."""
    n = x32.shape[0]
    a = np.empty((n, 2), np.float64)
    first = pc.dtype == np.float64
    for k in range(2):
        if first:
            xs = x32.astype(np.float64) * pc[k]                 # f32 @ f64 -> f64
        else:
            xs = (x32 * pc[k]).astype(np.float64)               # f32 @ f32 -> f32 product
        y = xs - np.float64(mu[k]) * np.float64(pc[k])
        q = (y * y).astype(np.float32)                          # log_prob buffer has X.dtype
        lp = np.float32(-0.5) * (np.float32(_LOG_2PI) + q)      # weak python scalars: f32 math
        if first:
            lp = lp.astype(np.float64) + np.log(pc[k])          # f32 array + f64 log-det
        else:
            lp = (lp + np.log(pc[k])).astype(np.float64)        # f32 log, f32 add
        a[:, k] = lp + np.log(w[k])
    return a


def _row_logsumexp(a):
    """sklearn/utils/_array_api.py:1338-1366 for two columns: the maximal entries are masked out,
    the rest is summed as exp(a - max), divided by the number m of maximal entries and fed to
    log1p; log(m) is added back."""
    amax = a.max(axis=1)
    is_max = a == amax[:, None]
    m = is_max.sum(axis=1).astype(np.float64)
    rest = np.where(is_max, -np.inf, a)
    with np.errstate(under="ignore"):
        s = np.exp(rest - amax[:, None]).sum(axis=1)
    s = np.where(s == 0, s, s / m)
    return np.log1p(s) + np.log(m) + amax


def _fit_numpy(x):
    x32 = np.ascontiguousarray(x, np.float32).reshape(-1)
    x64 = x32.astype(np.float64)
    w = np.array([0.5, 0.5], np.float64)
    mu = np.array([x32.min(), x32.max()], np.float64)
    pc = np.array([1.0, 1.0], np.float64)
    var = np.array([1.0, 1.0], np.float32)
    lower = -np.inf
    converged = False
    n_iter = 0
    for n_iter in range(1, MAX_ITER + 1):
        prev = lower
        a = _weighted_log_prob(x32, w, mu, pc)                  # E-step
        lpn = _row_logsumexp(a)
        with np.errstate(under="ignore"):
            resp = np.exp(a - lpn[:, None])
        nk = resp.sum(axis=0) + _EPS10                          # M-step
        mu = (resp.T @ x64) / nk
        var = np.empty(2, np.float32)
        for k in range(2):
            d = x64 - mu[k]
            var[k] = np.float32(((resp[:, k] * d) @ d) / nk[k])  # stored in an X.dtype buffer
            var[k] = var[k] + np.float32(REG_COVAR)              # in-place add on the f32 buffer
        w = nk / nk.sum()
        pc = (np.float32(1.0) / np.sqrt(var)).astype(np.float32)
        lower = lpn.mean()
        if abs(lower - prev) < TOL:
            converged = True
            break
    a = _weighted_log_prob(x32, w, mu, pc)                      # final E-step / predict / score
    return dict(weights=w, means=mu, variances=var.astype(np.float64), n_iter=n_iter,
                converged=converged, components=np.argmax(a, axis=1).astype(np.int64),
                scores=_row_logsumexp(a))


def fit_two_component(x, impl="sklearn"):
    """x: float32 [n,1] (or [n]) sorted ascending, n >= 2 (loss.py:189-192)."""
    x = np.ascontiguousarray(x, np.float32).reshape(-1, 1)
    if impl == "sklearn":
        return _fit_sklearn(x)
    if impl == "numpy":
        return _fit_numpy(x)
    raise ValueError(impl)


def positive_prefix_length(fit):
    """loss.py:206-217.  Foreground is mixture component 0 (whatever its mean).  With at least one
    foreground sample the positives are the sorted candidates up to and including the first
    foreground sample that attains the maximal foreground log-likelihood; without any, every
    candidate is positive."""
    comp = fit["components"]
    score = fit["scores"]
    fg = comp == 0
    if not fg.any():
        return int(comp.shape[0])
    best = score[fg].max()
    return int(np.nonzero(fg & (score == best))[0].min()) + 1


def component_margin(fit, x):
    """Smallest |a_0 - a_1| over the samples, a_k the weighted log-probabilities of the fitted components
    (float64, from the fitted parameters).  A value near 0 means `predict` (argmax, loss.py:202,206) is
    decided by rounding noise for that sample -- e.g. two candidates whose components collapse onto the
    same mean.  Tests exempt such a GT from mask-exactness (exemption (ii-b))."""
    x = np.asarray(x, np.float64).reshape(-1)
    w, mu, var = (np.asarray(fit[k], np.float64).reshape(2) for k in ("weights", "means", "variances"))
    a = [-0.5 * (np.log(2.0 * np.pi) + (x - mu[k]) ** 2 / var[k]) - 0.5 * np.log(var[k]) + np.log(w[k])
         for k in (0, 1)]
    return float(np.min(np.abs(a[0] - a[1])))


def structural_tie_margin(fit):
    """Gap between the two best foreground scores (inf if fewer than two foreground samples).
    Tests exempt a GT from mask-exactness when this is below 1e-5 (exemption (ii))."""
    fg = fit["components"] == 0
    s = np.sort(fit["scores"][fg])[::-1]
    return float(s[0] - s[1]) if s.shape[0] >= 2 else float("inf")
