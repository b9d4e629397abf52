"""BASELINE configs[1] at its full size (65 536 randomised LCO 1C discharges) through size-independent properties:
the oracle cannot run this many systems in a test, so the checks are invariants of the path itself --
bit-reproducibility, independence of a system's bits from its place in the batch and from its neighbours,
agreement of an embedded sample with the same systems run alone (which the oracle parity tests cover), and the
exact charge balance of a constant-current run."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

B = 65536


def _run(P, p, th, n_save_max=0):
    util.set_theta_batch(p, th)
    sol = P.simulate(p, I=-1, SOC=1, n_save_max=n_save_max)
    return sol.results[-1].summary.copy(), sol.Y.copy()


def test_full_size_batch_properties():
    import petlion_b200 as P
    p = P.petlion("LCO")
    tho = util.oracle_theta_batch(B)
    th = util.product_theta_from_oracle(p, tho)
    s1, Y1 = _run(P, p, th)
    # every system finished: SOC_min, V_min for the slow-diffusion draws, or a flagged per-system failure
    assert np.all(np.isin(s1["flag"], (1, 3, -2, -3)))
    ok = s1["flag"] >= 0
    assert np.mean(ok) > 0.995
    assert np.all(s1["n_steps"][ok] > 40) and np.all(s1["n_steps"][ok] < 200)
    # constant current: the trapezoid SOC update is exact, SOC_end = 1 - t_end/3600 (1C, interpolated end point included)
    np.testing.assert_allclose(s1["SOC_end"][ok], 1.0 - s1["t_end"][ok] / 3600.0, rtol=0, atol=2e-12)
    assert np.all(s1["I_end"][ok] == -1.0)
    assert np.all(s1["V_end"][ok] >= 2.5 - 1e-9) and np.all(s1["t_end"][ok] <= 3600.0 + 1e-6)
    hit_v = s1["flag"] == 1
    np.testing.assert_allclose(s1["V_end"][hit_v], 2.5, rtol=0, atol=1e-9)       # interpolated onto the bound
    # 1. bit-reproducible
    s2, Y2 = _run(P, p, th)
    assert s1.tobytes() == s2.tobytes() and np.array_equal(Y1, Y2, equal_nan=True)
    # 2. a system's bits do not depend on where it sits in the batch (work queue, CTA neighbours)
    perm = np.random.default_rng(5).permutation(B)
    s3, Y3 = _run(P, p, th[perm])
    assert s3.tobytes() == s1[perm].tobytes() and np.array_equal(Y3, Y1[perm], equal_nan=True)
    # 3. ... nor on the batch size: an embedded sample equals the same systems run alone
    idx = np.arange(0, B, B // 96)[:96]
    s4, Y4 = _run(P, p, th[idx])
    assert s4.tobytes() == s1[idx].tobytes() and np.array_equal(Y4, Y1[idx], equal_nan=True)
