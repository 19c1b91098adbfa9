"""numpy restatement of the sklearn call (oracle/gmm_oracle.py impl="numpy") against scikit-learn
itself on random candidate sets shaped like PAA's (2..45 sorted float32 losses)."""
import numpy as np

from oracle import gmm_oracle


def _random_sets(seed, count):
    rng = np.random.default_rng(seed)
    for _ in range(count):
        n = int(rng.integers(2, 46))
        k = int(rng.integers(0, n + 1))
        a = rng.normal(0.4, 0.15 * rng.random() + 0.01, k)
        b = rng.normal(1.5 + 2 * rng.random(), 0.5 * rng.random() + 0.01, n - k)
        yield np.sort(np.abs(np.concatenate([a, b])).astype(np.float32))


def test_numpy_restatement_equals_sklearn():
    for x in _random_sets(0, 300):
        f1 = gmm_oracle.fit_two_component(x, "sklearn")
        f2 = gmm_oracle.fit_two_component(x, "numpy")
        assert f1["n_iter"] == f2["n_iter"]
        assert np.array_equal(f1["components"], f2["components"])
        assert gmm_oracle.positive_prefix_length(f1) == gmm_oracle.positive_prefix_length(f2)
        np.testing.assert_allclose(f1["means"], f2["means"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(f1["weights"], f2["weights"], rtol=1e-10)
        np.testing.assert_allclose(f1["scores"], f2["scores"], rtol=1e-10, atol=1e-12)


def test_degenerate_sets():
    # two samples; all-equal samples; a far outlier
    for x in (np.array([0.5, 0.7], np.float32), np.full(9, 1.25, np.float32),
              np.array([0.1, 0.11, 0.12, 50.0], np.float32)):
        f1 = gmm_oracle.fit_two_component(x, "sklearn")
        f2 = gmm_oracle.fit_two_component(x, "numpy")
        assert f1["n_iter"] == f2["n_iter"]
        assert gmm_oracle.positive_prefix_length(f1) == gmm_oracle.positive_prefix_length(f2)


def test_prefix_rule():
    fit = dict(components=np.array([0, 0, 1, 0, 1]), scores=np.array([-1.0, -0.5, -0.1, -0.5, -3.0]))
    assert gmm_oracle.positive_prefix_length(fit) == 2     # first index attaining the fg maximum
    fit = dict(components=np.array([1, 1, 1]), scores=np.array([-1.0, -0.5, -0.1]))
    assert gmm_oracle.positive_prefix_length(fit) == 3     # no foreground -> all positive
