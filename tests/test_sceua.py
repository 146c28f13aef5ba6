"""Population-synchronous SCE-UA driver (caller side of R/Optim_GR2MSemiDistr.R:225-229 -> rtop::sceua)."""
import numpy as np

from gr2msemidistr_b200.sceua import sceua_batched


def rosenbrock(X):
    return np.sum(100.0 * (X[:, 1:] - X[:, :-1] ** 2) ** 2 + (1.0 - X[:, :-1]) ** 2, axis=1)


def test_default_path_converges_at_the_reference_budget():
    # defaults = what the reference gets from rtop::sceua: ngs = 5, maxn = Max.Functions = 5000 (Optim:49).
    # A wide population at that budget gets two shuffles and stops at f = 6..16 on this function.
    for seed in range(3):
        res = sceua_batched(rosenbrock, -2 * np.ones(4), -5 * np.ones(4), 5 * np.ones(4), seed=seed)
        assert res["value"] < 1e-3 and res["evaluations"] <= 5000 + 3 * 5
        assert np.allclose(res["par"], 1.0, atol=0.05)
        assert res["generations"] >= 20


def test_restarts_run_in_lockstep_and_return_the_best():
    calls = []

    def obj(X):
        calls.append(len(X))
        return rosenbrock(X)

    one = sceua_batched(rosenbrock, -2 * np.ones(8), -5 * np.ones(8), 5 * np.ones(8), maxn=3000, seed=[5, 0])
    res = sceua_batched(obj, -2 * np.ones(8), -5 * np.ones(8), 5 * np.ones(8), maxn=3000, restarts=6, seed=5)
    # restart 0 draws from the stream (seed, 0) and starts at x0: it is the single search, so best-of-6 <= it
    assert res["value"] <= one["value"]
    assert calls[0] == 6 * 5 * 17                     # all initial populations in one batch
    assert max(calls[1:]) <= 6 * 5 and 6 * 2990 <= res["evaluations"] <= 6 * (3000 + 15)
    assert sum(calls) == res["evaluations"]
    assert 0 <= res["restart"] < 6


def test_nan_objective_never_wins():
    def obj(X):
        f = rosenbrock(X)
        f[X[:, 0] < 0] = np.nan               # NA criterion (e.g. sd(sim) = 0): ranked last, like Inf
        return f
    res = sceua_batched(obj, 2 * np.ones(3), -5 * np.ones(3), 5 * np.ones(3), maxn=2000, seed=1)
    assert np.isfinite(res["value"]) and res["par"][0] >= 0
