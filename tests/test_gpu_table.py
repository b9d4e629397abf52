"""Tabulated time-varying inputs (the reference's run_function) on the GPU, through the C ABI,
against the CPU oracle and the printed results of examples/variable_input_functions.ipynb.

Tolerances: as tests/test_gpu_parity.py (identical step sequences; t, V, SOC within rtol 1e-6).  The runs
cross a discontinuity through dozens of failed/retried steps, so the step-sequence comparison here is a
strong check of the failure paths of the integrator (error-test failures, IDA_ERR_FAIL returns that are
not errors for a run_function, the re-initialisation of checks.jl:341-364).
"""
import math

import numpy as np
import pytest

import oracle as O
from tests import util
from tests.test_gpu_parity import _compare_runs

pytestmark = pytest.mark.gpu

STEP = ([0.0, 100.0, 100.0], [1.0, 1.0, 0.5])        # I_fun1(t) = t < 100 ? 1 : 0.5


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


@pytest.fixture(scope="module")
def lco(P):
    return P.petlion("LCO")


def _nominal(P, lco):
    for k, v in zip(lco.θ_keys, P.petlion("LCO").θ.values()):
        lco.θ[k] = v


@pytest.mark.parametrize("tdiscon", [(), (100.0,)])
def test_step_current_nominal(P, lco, goldens, tdiscon):
    """variable_input_functions.ipynb cells 6 and 8: 1C -> C/2 at t = 100 s, with and without tdiscon"""
    _nominal(P, lco)
    m = O.make_model("LCO")
    sol = P.simulate(lco, 200, I=P.Table(*STEP), SOC=0, tdiscon=list(tdiscon))
    ref = O.simulate_batch(m, O.theta_defaults("LCO"), O.make_run("I", tf=200, table=STEP, tdiscon=tdiscon),
                           O.default_opts(), O.default_bounds("LCO"), SOC0=0.0, n_save_max=512)
    s = sol.results[-1].summary
    assert s["flag"][0] == 0 and s["t_end"][0] == 200.0
    assert s["n_reinit"][0] == ref["n_reinit"][0] == (0 if tdiscon else 1)
    # approaching an unannounced jump, whether a trial step crosses it is decided at the 1e-9 s level, so the
    # count of failed attempts may differ by one or two between two correct implementations
    assert abs(int(s["n_netf"][0]) - int(ref["n_netf"][0])) <= (0 if tdiscon else 2) and s["n_netf"][0] >= 5
    _compare_runs(sol, ref)
    n = sol.n_points[0]
    np.testing.assert_allclose(sol.I[0, :n], ref["traj"]["I"][0, :n], rtol=1e-9, atol=1e-12)
    if tdiscon:   # the integrator stops at tdiscon - reltol/2 (model_evaluation.jl:295-297)
        assert np.any(np.abs(sol.t[0, :n] - (100.0 - 0.5e-3)) < 1e-9)
    g = goldens["function_inputs"]["step_tdiscon" if tdiscon else "step"]
    I1C = lco.I1C()[0]
    assert abs(s["V_end"][0] - g["V"]) < 1.5e-4
    assert abs(s["I_end"][0] * I1C * s["V_end"][0] - g["P"]) < 5e-5 * g["P"]
    assert abs(s["SOC_end"][0] - g["SOC"]) < 5e-5


@pytest.mark.parametrize("name", ["ramp_100", "ramp_10"])
def test_current_ramps(P, lco, goldens, name):
    """variable_input_functions.ipynb cells 12 and 14: I(t) = ramp_val * t"""
    _nominal(P, lco)
    g = goldens["function_inputs"][name]
    tab = ([0.0, 100.0], [0.0, 100.0 * g["ramp_val"]])
    sol = P.simulate(lco, 100, I=P.Table(*tab), SOC=0)
    ref = O.simulate_batch(O.make_model("LCO"), O.theta_defaults("LCO"), O.make_run("I", tf=100, table=tab),
                           O.default_opts(), O.default_bounds("LCO"), SOC0=0.0, n_save_max=512)
    _compare_runs(sol, ref)
    s = sol.results[-1].summary
    I1C = lco.I1C()[0]
    assert abs(s["V_end"][0] - g["V"]) < 1.5e-4
    assert abs(s["I_end"][0] * I1C * s["V_end"][0] - g["P"]) < 5e-5 * g["P"]
    assert abs(s["SOC_end"][0] - g["SOC"]) < 5e-5


def test_step_current_randomised_batch_with_scales(P, lco):
    """randomised parameters and a per-system scale on the profile; a discontinuity the integrator is not told about"""
    m = O.make_model("LCO")
    B = 96
    tho = util.oracle_theta_batch(B, first=300)
    util.set_theta_batch(lco, util.product_theta_from_oracle(lco, tho))
    scale = 0.5 + 1.5 * util.splitmix_u01(7, np.arange(B), 3)
    tab = ([0.0, 60.0, 60.0, 150.0, 150.0], [1.0, 1.0, -0.5, -0.5, 0.25])
    sol = P.simulate(lco, 240, I=P.Table(*tab, scale=scale), SOC=0.3)
    ref = O.simulate_batch(m, tho, O.make_run("I", tf=240, table=tab), O.default_opts(), O.default_bounds("LCO"),
                           SOC0=0.3, values=scale, n_save_max=512, nthreads=8)
    s = sol.results[-1].summary
    assert np.all(ref["flag"] == 0) and np.all(ref["n_reinit"] >= 1)
    same = s["n_steps"] == ref["n_steps"]
    assert np.array_equal(s["n_reinit"][same], ref["n_reinit"][same])
    # (two runs that step over an untold jump of the input differently differ by more than 5 reltol next to it)
    _compare_runs(sol, ref, min_identical=0.9, flip_tol=2e-2)
    np.testing.assert_allclose(s["I_end"], 0.25 * scale, rtol=1e-9)


@pytest.mark.parametrize("method,func,soc", [("P", lambda t: 40.0 * math.sin(t), 0.5),
                                             ("V", lambda t: 3.9 + 0.05 * math.cos(t), 0.5)])
def test_power_and_voltage_functions(P, lco, method, func, soc):
    """variable_input_functions.ipynb cells 17-19 (sinusoidal P(t), V(t)), tabulated at 0.05 s"""
    m = O.make_model("LCO")
    B = 16
    tho = util.oracle_theta_batch(B, first=900)
    util.set_theta_batch(lco, util.product_theta_from_oracle(lco, tho))
    tab = P.Table.sample(func, np.linspace(0.0, 10.0, 201))
    sol = P.simulate(lco, 10, SOC=soc, **{method: tab})
    ref = O.simulate_batch(m, tho, O.make_run(method, tf=10, table=(tab.t, tab.v)), O.default_opts(),
                           O.default_bounds("LCO"), SOC0=soc, n_save_max=512, nthreads=8)
    assert np.all(ref["flag"] == 0)
    _compare_runs(sol, ref, min_identical=0.8)
    s = sol.results[-1].summary
    if method == "V":
        np.testing.assert_allclose(s["V_end"], func(10.0), rtol=2e-5)    # Newton tolerance of the step at reltol 1e-3
    else:
        np.testing.assert_allclose(s["I_end"] * lco.I1C(B) * s["V_end"], func(10.0), rtol=2e-3)


def test_table_after_constant_current_segment(P, lco):
    """simulate! with a function input: its times restart at 0 and the continuation stop at t = 1 merges with tdiscon"""
    m = O.make_model("LCO")
    B = 12
    tho = util.oracle_theta_batch(B, first=40)
    util.set_theta_batch(lco, util.product_theta_from_oracle(lco, tho))
    sol = P.simulate(lco, 300, I=1, SOC=0.2)
    ref = O.simulate_batch(m, tho, O.make_run("I", 1.0, tf=300), O.default_opts(), O.default_bounds("LCO"), SOC0=0.2,
                           n_save_max=512, nthreads=8)
    _compare_runs(sol, ref)
    tab = ([0.0, 0.5, 0.5, 30.0, 30.0], [1.0, 1.0, 2.0, 2.0, -1.0])
    P.simulate_(sol, lco, 90, I=P.Table(*tab), tdiscon=[0.5, 30.0])
    ref2 = O.simulate_batch(m, tho, O.make_run("I", tf=90, table=tab, tdiscon=[0.5, 30.0], new_run=False),
                            O.default_opts(), O.default_bounds("LCO"), state=ref["state"], n_save_max=512, nthreads=8)
    s2 = sol.results[-1].summary
    same = s2["n_steps"] == ref2["n_steps"]
    assert np.mean(same) >= 0.75
    assert np.array_equal(s2["flag"][same], ref2["flag"][same]) and np.all(ref2["flag"] == 0)
    np.testing.assert_allclose(s2["t_end"], 390.0, rtol=1e-12)
    np.testing.assert_allclose(s2["V_end"][same], ref2["V_end"][same], rtol=1e-6)
    np.testing.assert_allclose(s2["SOC_end"][same], ref2["SOC_end"][same], rtol=1e-6)
    np.testing.assert_allclose(s2["I_end"], -1.0, rtol=1e-9)


def test_constant_table_equals_constant_run(P, lco):
    """test/runtests.jl:35 -- `I = (t) -> 1` gives the same result as `I = 1`"""
    _nominal(P, lco)
    a = P.simulate(lco, 500, I=1, SOC=0)
    b = P.simulate(lco, 500, I=P.Table([0.0, 1000.0], [1.0, 1.0]), SOC=0)
    n = a.n_points[0]
    assert b.n_points[0] == n
    assert np.array_equal(a.t[0, :n], b.t[0, :n]) and np.array_equal(a.V[0, :n], b.V[0, :n])


def test_table_argument_errors(P, lco):
    _nominal(P, lco)
    with pytest.raises(ValueError):
        P.Table([0.0, 2.0, 1.0], [1.0, 1.0, 1.0])
    with pytest.raises(NotImplementedError):
        P.simulate(lco, 10, I=lambda t: 1.0, SOC=0)
    with pytest.raises(RuntimeError, match="more than two knots"):
        P.simulate(lco, 10, I=P.Table([0.0, 1.0, 1.0, 1.0], [1.0, 1.0, 2.0, 3.0]), SOC=0)
    with pytest.raises(ValueError, match="dT takes"):
        P.simulate(P.petlion("LCO", temperature=True), 10, dT=P.Table([0.0, 1.0], [0.0, 0.0]), SOC=0)


@pytest.mark.parametrize("family", ["thermal", "sei", "wide"])
def test_table_and_state_rows_in_the_other_families(P, family):
    """the extended kernel (tabulated input + kept state rows) of every compiled family against the oracle"""
    kw = dict(thermal=dict(temperature=True), sei=dict(aging="SEI"), wide=dict(N_p=20, N_s=20, N_n=20))[family]
    okw = dict(thermal=dict(temperature=True), sei=dict(aging=True), wide=dict(N_p=20, N_s=20, N_n=20))[family]
    p = P.petlion("LCO", **kw)
    m = O.make_model("LCO", **okw)
    B = 6
    tho = util.oracle_theta_batch(B, first=77)
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    tab = ([0.0, 50.0, 50.0, 120.0], [1.0, 1.0, 2.0, 0.5])          # a jump, then a ramp down
    sol = P.simulate(p, 120, I=P.Table(*tab), SOC=0.2, tdiscon=[50.0], outputs="all")
    ref = O.simulate_batch(m, tho, O.make_run("I", tf=120, table=tab, tdiscon=[50.0]), O.default_opts(),
                           O.default_bounds("LCO"), SOC0=0.2, n_save_max=512, nthreads=6)
    assert np.all(ref["flag"] == 0)
    s = sol.results[-1].summary
    same = s["n_steps"] == ref["n_steps"]
    assert np.mean(same) >= 0.6
    np.testing.assert_allclose(s["V_end"][same], ref["V_end"][same], rtol=1e-6)
    np.testing.assert_allclose(s["SOC_end"], ref["SOC_end"], rtol=1e-5)
    np.testing.assert_allclose(s["I_end"], 0.5, rtol=1e-9)
    for k in range(B):
        n = sol.n_points[k]
        ps = sol.states[k, :n, p.ind["Φ_s"]]
        np.testing.assert_allclose(ps[:, 0] - ps[:, -1], sol.V[k, :n], rtol=1e-13)
        np.testing.assert_array_equal(sol.states[k, :n, p.ind["I"]][:, 0], sol.I[k, :n])
        np.testing.assert_array_equal(sol.states[k, n - 1], sol.Y[k])
        assert np.all(np.isfinite(sol.states[k, :n]))


def test_user_stop_times(P, lco):
    """opts.tstops (params.jl:272): the integrator lands exactly on the requested times; on a continuation they
    merge with the stop at t = 1 (model_evaluation.jl:288-310)"""
    m = O.make_model("LCO")
    B = 10
    tho = util.oracle_theta_batch(B, first=11)
    util.set_theta_batch(lco, util.product_theta_from_oracle(lco, tho))
    ts = [0.5, 100.0, 250.5, 5000.0]
    sol = P.simulate(lco, 1000, I=-1, SOC=1, tstops=ts)
    ref = O.simulate_batch(m, tho, O.make_run("I", -1.0, tf=1000, tstops=ts), O.default_opts(), O.default_bounds("LCO"),
                           SOC0=1.0, n_save_max=512, nthreads=8)
    _compare_runs(sol, ref, min_identical=0.9)
    for k in range(B):
        t = sol.t[k, :sol.n_points[k]]
        assert all(np.any(t == x) for x in (0.5, 100.0, 250.5)) and t[-1] == 1000.0
    P.simulate_(sol, lco, 300, I=1, tstops=[0.25, 40.0])
    ref2 = O.simulate_batch(m, tho, O.make_run("I", 1.0, tf=300, tstops=[0.25, 40.0], new_run=False), O.default_opts(),
                            O.default_bounds("LCO"), state=ref["state"], n_save_max=512, nthreads=8)
    s2 = sol.results[-1].summary
    same = s2["n_steps"] == ref2["n_steps"]
    assert np.mean(same) >= 0.8
    np.testing.assert_allclose(s2["V_end"][same], ref2["V_end"][same], rtol=1e-6)
    for k in range(B):
        t = sol.t[k, :sol.n_points[k]] - 1000.0
        assert all(np.any(np.abs(t - x) < 1e-9) for x in (0.25, 1.0, 40.0))
    # the stop list is a model option: cleared again for the following runs
    a = P.simulate(lco, 500, I=-1, SOC=1)
    assert not np.any(a.t[0, :a.n_points[0]] == 100.0)
