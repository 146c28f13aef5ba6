"""R >= 4.0 round(x, digits) restatement -- spot values and properties."""
import numpy as np

from oracle import twin


def test_spot_values(oracle):
    # not rint(x*10)/10
    assert oracle.r_round(0.15, 1) == 0.1
    assert oracle.r_round(0.25, 1) == 0.2
    assert oracle.r_round(0.35, 1) == 0.3
    assert oracle.r_round(29.65, 1) == 29.6      # 1.186 * 25.0, the bundled example's set 0
    assert oracle.r_round(2.5, 0) == 2.0
    assert oracle.r_round(-0.15, 1) == -0.1
    assert oracle.r_round(0.0, 1) == 0.0
    assert oracle.r_round(2.675, 2) == 2.67       # 2.675 is below the tie in binary
    assert oracle.r_round(1e-9, 1) == 0.0
    assert np.isnan(oracle.r_round(np.nan, 1))
    assert oracle.r_round(np.inf, 2) == np.inf
    assert oracle.r_round(123456789012.345, 3) == 123456789012.345
    assert oracle.r_round(1234.5678, 3) == 1234.568


def test_result_is_a_decimal_grid_point_and_nearest(oracle, rng):
    x = rng.uniform(0, 4000, 20000)
    for d in (1, 2, 3):
        r = oracle.r_round(x, d)
        # r is the double nearest to some k/10^d, and |r - x| <= half a quantum (+1 ulp slack)
        k = np.rint(r * 10 ** d)
        assert np.array_equal(r, k / 10 ** d)
        assert np.all(np.abs(r - x) <= 0.5 / 10 ** d + 1e-12)


def test_already_rounded_is_fixed_point(oracle, rng):
    k = rng.integers(0, 40000, 20000)
    x = k / 10.0
    assert np.array_equal(oracle.r_round(x, 1), x)
    assert np.array_equal(oracle.r_round(x, 2), x)


def test_matches_numpy_twin(oracle, rng):
    x = np.concatenate([rng.uniform(-500, 500, 50000), rng.integers(0, 10000, 20000) / 20.0,
                        1.186 * (rng.integers(0, 5000, 20000) / 10.0)])
    for d in (1, 2, 3):
        assert np.array_equal(oracle.r_round(x, d), twin.r_round(x, d))
