"""GPU tests of the four stop predicates round 1 left untested -- c_s_n_max, c_e_min, eta_plating_min, dfilm_max
(src/checks.jl:141-224) -- with their t_frac back-interpolation (model_evaluation.jl:369-382), against the oracle
(tests/test_oracle_stops.py pins the oracle's own behaviour); plus the hand-over of hard failures between the
segments of a protocol."""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


def _both(P, p, m, tho, I, soc, **bo):
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    sol = P.simulate(p, I=I, SOC=soc, **bo)
    b = O.default_bounds("LCO", **{("eta_plating_min" if k == "η_plating_min" else k): v for k, v in bo.items()})
    ref = O.simulate_batch(m, tho, O.make_run("I", I), O.default_opts(), b, SOC0=soc, nthreads=8, n_save_max=512)
    s = sol.results[-1].summary
    same = s["n_steps"] == ref["n_steps"]
    assert same.mean() >= 0.8
    assert np.array_equal(s["flag"][same], ref["flag"][same])
    np.testing.assert_allclose(s["t_end"][same], ref["t_end"][same], rtol=1e-6)
    np.testing.assert_allclose(s["V_end"][same], ref["V_end"][same], rtol=1e-6)
    # systems on a different step sequence still exit on the same predicate, at the integration tolerance
    np.testing.assert_allclose(s["t_end"], ref["t_end"], rtol=5e-3)
    return sol, ref, s


def test_stop_c_s_n_max(P):
    p, m = P.petlion("LCO"), O.make_model("LCO")
    tho = util.oracle_theta_batch(32, first=40)
    sol, ref, s = _both(P, p, m, tho, 2.0, 0.0, c_s_n_max=0.5)
    assert np.all(s["flag"] == 6) and np.all(ref["flag"] == 6)
    surf = sol.Y[:, p.ind["c_s_avg"]][:, p.N.p * p.N.r_p:][:, p.N.r_n - 1::p.N.r_n]
    cmax = tho[:, O.theta_names().index("c_max_n")]
    np.testing.assert_allclose(surf.max(axis=1) / cmax, 0.5, rtol=1e-12)      # the blended end state sits on the bound
    assert sol.t[0, sol.n_points[0] - 1] == s["t_end"][0]                      # ... and is the last saved row


def test_stop_c_e_min(P):
    p, m = P.petlion("LCO"), O.make_model("LCO")
    tho = util.oracle_theta_batch(32, first=80)
    sol, ref, s = _both(P, p, m, tho, 2.0, 0.0, c_e_min=800.0)
    assert np.all(s["flag"] == 9) and np.all(ref["flag"] == 9)
    np.testing.assert_allclose(sol.Y[:, p.ind["c_e"]].min(axis=1), 800.0, rtol=1e-6)


def test_stop_eta_plating_min(P):
    p, m = P.petlion("LCO"), O.make_model("LCO")
    tho = util.oracle_theta_batch(32, first=120)
    sol, ref, s = _both(P, p, m, tho, 4.0, 0.0, η_plating_min=0.05)
    assert np.all(s["flag"] == 11) and np.all(ref["flag"] == 11)
    eta = sol.Y[:, p.ind["Φ_s"]][:, p.N.p] - sol.Y[:, p.ind["Φ_e"]][:, p.N.p + p.N.s]
    np.testing.assert_allclose(eta, 0.05, rtol=1e-10)


@pytest.mark.parametrize("grid", [{}, dict(N_p=20, N_s=20, N_n=20)])
def test_stop_dfilm_max(P, grid):
    p, m = P.petlion("LCO", aging="SEI", **grid), O.make_model("LCO", aging=True, **grid)
    tho = util.oracle_theta_batch(16, first=160)
    sol, ref, s = _both(P, p, m, tho, 1.0, 0.0, V_max=4.2, dfilm_max=3e-15)
    assert np.all(s["flag"] == 10) and np.all(ref["flag"] == 10)
    names = O.theta_names()
    rate = (-sol.Y[:, p.ind["j_s"]] * (tho[:, names.index("M_n")] / tho[:, names.index("rho_n")])[:, None]).max(axis=1)
    np.testing.assert_allclose(rate, 3e-15, rtol=2e-2)


def test_two_bounds_smallest_fraction_wins(P):
    p, m = P.petlion("LCO"), O.make_model("LCO")
    tho = util.oracle_theta_batch(32, first=200)
    sol, ref, s = _both(P, p, m, tho, 4.0, 0.0, η_plating_min=0.05, c_e_min=300.0)
    assert set(np.unique(s["flag"])) <= {9, 11}


def test_hard_failure_is_carried_across_segments(P):
    """a system that failed hard in one segment must not come back to life in the next one:
    state_t = NaN marks it, simulate!() passes it through with flag -7, its neighbours are untouched"""
    p = P.petlion("LCO")
    B = 16
    tho = util.oracle_theta_batch(B, first=240)
    th = util.product_theta_from_oracle(p, tho)
    bad = [3, 11]
    th_bad = th.copy()
    th_bad[bad, p.θ_keys.index("D_sp")] = np.nan
    util.set_theta_batch(p, th_bad)
    sol = P.simulate(p, 600.0, I=1, SOC=0.2)
    f1 = sol.results[-1].summary["flag"]
    assert np.all(f1[bad] < 0) and np.all(np.delete(f1, bad) == 0)
    assert np.all(np.isnan(sol._t_end[bad]))
    P.simulate_(sol, p, 600.0, I="rest")
    f2 = sol.results[-1].summary
    assert np.all(f2["flag"][bad] == -7) and np.all(np.delete(f2["flag"], bad) == 0)
    assert np.all(f2["n_steps"][bad] == 0)
    # the healthy systems equal a run that never had the poisoned neighbours
    util.set_theta_batch(p, th)
    good = P.simulate(p, 600.0, I=1, SOC=0.2)
    P.simulate_(good, p, 600.0, I="rest")
    keep = np.delete(np.arange(B), bad)
    assert np.array_equal(good.results[-1].summary["V_end"][keep], f2["V_end"][keep])
    # the oracle does the same
    m = O.make_model("LCO")
    tho_bad = tho.copy(); tho_bad[bad, O.theta_names().index("D_sp")] = np.nan
    r1 = O.simulate_batch(m, tho_bad, O.make_run("I", 1.0, tf=600.0), O.default_opts(), O.default_bounds("LCO"), SOC0=0.2, nthreads=4)
    r2 = O.simulate_batch(m, tho_bad, O.make_run("I", 0.0, tf=600.0, input_kind="rest", new_run=False), O.default_opts(),
                          O.default_bounds("LCO"), state=r1["state"], nthreads=4)
    assert np.all(r2["flag"][bad] == -7) and np.all(np.delete(r2["flag"], bad) == 0)


def test_two_handles_on_two_devices(P):
    """every ABI call runs on its handle's device, whatever the current device is"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    a, b = P.petlion("LCO", device=0), P.petlion("LCO", device=1)
    ref = P.simulate(a, I=-1, SOC=1).results[-1].summary
    torch.cuda.set_device(1)
    ra = P.simulate(a, I=-1, SOC=1).results[-1].summary
    torch.cuda.set_device(0)
    rb = P.simulate(b, I=-1, SOC=1).results[-1].summary
    assert torch.cuda.current_device() == 0
    for r in (ra, rb):
        assert r["n_steps"][0] == ref["n_steps"][0] and r["V_end"][0] == ref["V_end"][0]
    Y0 = b.initial_guess(np.array([0.5]))
    res, nz = b.resjac(Y0, np.zeros_like(Y0), 0.1)
    res0, nz0 = a.resjac(Y0, np.zeros_like(Y0), 0.1)
    assert np.array_equal(res, res0) and np.array_equal(nz, nz0)
