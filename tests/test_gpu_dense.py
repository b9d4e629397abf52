"""Dense output (simulate(p, tf::AbstractVector), src/model_evaluation.jl:13, 80, 148-149): the device evaluates the
BDF interpolant of the covering step at the requested times (plb_set_dense_output); parity vs the oracle's
interpolant on the same seeded inputs.  sol(t): the reference's post-hoc spline (save_outputs.jl:74-133)."""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


def test_dense_rows_equal_oracle_interpolant(P):
    p = P.petlion("LCO")
    m = O.make_model("LCO")
    B = 48
    tho = util.oracle_theta_batch(B, first=300)
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    td = np.concatenate([np.arange(0.0, 3590.0, 50.0), [1e6]])      # (not 3600 s: the runs end there to round-off)
    sol = P.simulate(p, td, I=-1, SOC=1, outputs="all")
    ref = O.simulate_batch(m, tho, O.make_run("I", -1.0), O.default_opts(), O.default_bounds("LCO"), SOC0=1.0,
                           nthreads=8, dense_t=td, dense_Y=True)
    s = sol.results[-1].summary
    # the same decisions: EVERY counter agrees (equal step counts alone can hide an error-test failure on one side and a
    # convergence failure on the other: then the step times differ and so do the rows, at the integrator's tolerance)
    same = np.ones(B, dtype=bool)
    for c in ("flag", "n_steps", "n_res", "n_jac", "n_netf", "n_ncfn"):
        same &= s[c] == ref[c]
    assert same.mean() > 0.9
    d, r = sol.dense, ref["dense"]
    assert np.array_equal(d["n"][same], r["n"][same])
    assert d["n"].max() < td.size          # the row at 1e6 s is past every run's end
    for i in np.where(same)[0]:
        n = d["n"][i]
        np.testing.assert_allclose(d["V"][i, :n], r["V"][i, :n], rtol=1e-6)
        np.testing.assert_allclose(d["I"][i, :n], r["I"][i, :n], rtol=1e-9)
        np.testing.assert_allclose(d["SOC"][i, :n], r["SOC"][i, :n], rtol=0, atol=1e-9)
        scale = np.maximum(np.abs(r["Y"][i, :n]), 1e-6 * np.abs(r["Y"][i, :n]).max(axis=0, keepdims=True) + 1e-30)
        assert np.max(np.abs(d["Y"][i, :n] - r["Y"][i, :n]) / scale) < 1e-5
        assert np.all(np.isnan(d["V"][i, n:]))
    # the first requested time is the start: exactly the first saved row
    assert np.array_equal(d["V"][:, 0], sol.V[:, 0]) and np.array_equal(d["SOC"][:, 0], np.ones(B))
    # a requested time that IS a step time reproduces the saved row
    i = int(np.where(same)[0][0])
    tq = sol.t[i, 5:9]
    p1 = P.petlion("LCO"); util.set_theta_batch(p1, util.product_theta_from_oracle(p1, tho[i:i + 1]))
    s1 = P.simulate(p1, np.concatenate([tq, [1e6]]), I=-1, SOC=1)
    np.testing.assert_allclose(s1.dense["V"][0, :4], sol.V[i, 5:9], rtol=1e-12)


def test_dense_thermal_continuation(P):
    """CC to 4.1 V then V = :hold: each segment fills the requested (global) times it covers"""
    W = util.PROTOCOLS["cfg3i"]
    p = P.petlion("LCO", temperature=True)
    B = 16
    tho = util.oracle_theta_batch(B, first=700)
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    td = np.arange(0.0, 3000.0, 20.0)
    ref = util.oracle_protocol(W, tho, O.default_opts(), dense_t=td)
    sol, dense = util.gpu_protocol(P, p, W, dense_t=td)
    for k in range(2):
        s = sol.results[k].summary
        # "same decisions" = every counter of this and the earlier segment agrees (step counts alone are too weak: two
        # runs with equal counts can differ in one Newton iteration and sit 1e-5 apart in the algebraic current)
        same = np.ones(s["flag"].size, dtype=bool)
        for kk in range(k + 1):
            for c in ("flag", "n_steps", "n_res", "n_jac", "n_netf", "n_ncfn"):
                same &= sol.results[kk].summary[c] == ref[kk][c]
        assert same.mean() >= 0.7
        d, r = dense[k], ref[k]["dense"]
        for i in np.where(same)[0]:
            fill = ~np.isnan(r["V"][i])
            assert np.array_equal(fill, ~np.isnan(d["V"][i]))
            np.testing.assert_allclose(d["V"][i, fill], r["V"][i, fill], rtol=1e-6)
            np.testing.assert_allclose(d["T"][i, fill], r["T"][i, fill], rtol=1e-6)
            np.testing.assert_allclose(d["I"][i, fill], r["I"][i, fill], rtol=3e-5, atol=1e-8)
    g = util.merge_dense(dense)
    # no requested time before the end of the protocol is left open
    t_end = sol.results[-1].summary["t_end"]
    for i in range(B):
        assert not np.isnan(g["V"][i, td < t_end[i] - 1e-6]).any()


def test_sol_call_is_the_reference_spline(P):
    p = P.petlion("LCO")
    sol = P.simulate(p, I=-1, SOC=1)
    n = sol.n_points[0]
    # at the saved times an interpolating spline returns the saved rows
    out = sol(sol.t[0, :n])
    np.testing.assert_allclose(out["V"][0], sol.V[0, :n], rtol=1e-12)
    # between them it is a cubic through the neighbours: close to the integrator's own interpolant
    tq = np.arange(10.0, 3590.0, 37.0)
    dn = P.simulate(p, np.concatenate([tq, [1e6]]), I=-1, SOC=1).dense
    np.testing.assert_allclose(sol(tq)["V"][0], dn["V"][0, :tq.size], rtol=2e-3)
    # outside the run: nearest value ("interpolate") or the extrapolated polynomial
    assert sol(1e5)["V"][0, 0] == pytest.approx(sol.V[0, n - 1])
    assert sol(1e5, interp_bc="extrapolate")["V"][0, 0] != pytest.approx(sol.V[0, n - 1])
    with pytest.raises(ValueError):
        sol(1.0, interp_bc="x")


def test_dense_request_is_one_shot_and_validated(P):
    import ctypes as C
    from petlion_b200 import _lib
    p = P.petlion("LCO")
    L = _lib.lib()
    bad = np.array([3.0, 2.0])
    assert L.plb_set_dense_output(p._h, 2, bad.ctypes.data, None, None, None, None, None, None, 0) != 0
    sol = P.simulate(p, np.array([0.0, 100.0, 1e6]), I=-1, SOC=1)
    assert sol.dense["n"][0] == 2
    sol2 = P.simulate(p, I=-1, SOC=1)
    assert sol2.dense is None and sol2.results[-1].summary["n_steps"][0] == 80


def test_truncated_trajectories_are_flagged(P):
    p = P.petlion("LCO")
    with pytest.warns(RuntimeWarning, match="n_save_max"):
        sol = P.simulate(p, I=-1, SOC=1, n_save_max=40)
    assert sol.truncated[0] and sol.n_points[0] == 40 and sol.results[-1].summary["n_steps"][0] == 80
    sol = P.simulate(p, I=-1, SOC=1)
    assert not sol.truncated[0]
