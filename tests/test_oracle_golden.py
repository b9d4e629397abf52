"""CPU tests: the oracle against the reference's own known answers (executed-notebook outputs,
tests/golden/reference_goldens.json, extracted by tests/golden/make_golden.py)."""
import numpy as np
import pytest

import oracle as O


def _names():
    return O.theta_names()


def test_theta_defaults_match_reference_dict(goldens):
    # examples/updating_parameters.ipynb cell 3 prints the full LCO dict
    ref = goldens["theta_LCO"]
    uni = {"T0": "T₀", "c_e0": "c_e₀", "t_plus": "t₊"}
    greek = {"theta_": "θ_", "sigma_": "σ_", "eps_": "ϵ_", "lambda_": "λ_", "rho_": "ρ_"}
    th = O.theta_dict("LCO")
    checked = 0
    for k, v in th.items():
        rk = uni.get(k, k)
        for a, b in greek.items():
            if rk.startswith(a):
                rk = b + rk[len(a):]
        if rk in ref:
            assert v == pytest.approx(ref[rk], rel=1e-15), k
            checked += 1
    assert checked >= 55


def test_I1C_exact(goldens):
    assert O.calc_I1C(O.theta_defaults("LCO")) == goldens["theta_LCO"]["I1C"]


def test_layout_and_pattern_sizes():
    m = O.make_model("LCO")
    L = O.layout(m)
    assert (L.N_diff, L.N_alg, L.N_tot) == (230, 71, 301)
    cp, rv = O.jac_pattern(m, "I")
    assert len(rv) == 2139                                          # SURVEY §8, [probe]
    cp, rv = O.jac_pattern(m, "V")
    assert len(rv) == 2140
    # row-group counts quoted in SURVEY App. A
    cp, rv = O.jac_pattern(m, "I")
    rows = np.asarray(rv)
    assert np.sum(rows < 30) == 108 and np.sum((rows >= 30) & (rows < 230)) == 1660
    assert np.sum((rows >= 230) & (rows < 250)) == 100 and np.sum((rows >= 250) & (rows < 280)) == 192
    assert np.sum((rows >= 280) & (rows < 300)) == 78 and np.sum(rows == 300) == 1


def test_newton_init_V0_of_2C_charge(goldens):
    # examples/model_inputs_and_outputs.ipynb: sol.V[1] of simulate(p, I=2, SOC=0, V_max=4.1)
    m = O.make_model("LCO"); th = O.theta_defaults("LCO"); L = O.layout(m)
    y0 = O.initial_guess(m, th, 0.0); y0[L.I] = 2.0
    it, y, yp = O.newton_init(m, th, O.make_run("I", 2.0), O.default_opts(), y0)
    assert it == 4
    V0 = y[L.phi_s] - y[L.phi_s + 19]
    assert abs(V0 - goldens["V_2C_charge"]["head13"][0]) < 5e-11


@pytest.mark.parametrize("eps_p", ["0.385", "0.485", "0.585"])
def test_ida_step_ladder_matches_notebook(goldens, eps_p):
    """The three 1C discharges of examples/updating_parameters.ipynb cell 5: every IDA step time
    (decoded from the SVG polyline) must be reproduced, with the same number of steps."""
    m = O.make_model("LCO"); names = _names()
    th = O.theta_defaults("LCO"); th[names.index("eps_p")] = float(eps_p)
    r = O.simulate_batch(m, th, O.make_run("I", -1.0), O.default_opts(), O.default_bounds("LCO"),
                         SOC0=1.0, n_save_max=400)
    g = goldens["ladder_1C_discharge"]["eps_p"][eps_p]
    gt, gV = np.array(g["t"]), np.array(g["V"])
    n = r["traj_n"][0]
    t, V = r["traj"]["t"][0, :n], r["traj"]["V"][0, :n]
    assert n == len(gt)                                   # 81 / 66 / 55 points
    assert r["flag"][0] == 3                              # "Below min. SOC"
    assert abs(t[-1] - 3600.0) < 1e-6
    # pixel resolution: ~2 ms absolute on small steps, ~1e-3 relative on large ones
    assert np.all(np.abs(t - gt) <= 0.004 + 2e-3 * gt)
    # voltage decoded to ~1e-4 V (+ tick-label calibration)
    assert np.max(np.abs(V - gV)) < 1.5e-3
    assert abs(V[-1] - gV[-1]) < 2e-4


def test_printed_summary_1C(goldens):
    m = O.make_model("LCO"); th = O.theta_defaults("LCO")
    r = O.simulate_batch(m, th, O.make_run("I", -1.0), O.default_opts(), O.default_bounds("LCO"), SOC0=1.0)
    s = goldens["summaries"]["1C_discharge"]
    assert r["t_end"][0] == pytest.approx(s["t_s"], abs=1e-6)
    assert r["V_end"][0] == pytest.approx(s["V"], abs=2e-3)      # printed by an older PETLION version
    assert abs(r["SOC_end"][0]) < 1e-9


def test_cccv_protocol_close_to_older_notebook(goldens):
    """CC-CV notebook was executed with an older PETLION: indicative 3-digit agreement only."""
    m = O.make_model("LCO"); th = O.theta_defaults("LCO")
    b = O.default_bounds("LCO", V_max=4.1)
    r = O.simulate_batch(m, th, O.make_run("I", 2.0), O.default_opts(), b, SOC0=0.0)
    s = goldens["summaries"]["2C_CC_to_4.1V"]
    assert r["flag"][0] == 2
    assert r["t_end"][0] == pytest.approx(s["t_s"], rel=1e-3)
    assert r["SOC_end"][0] == pytest.approx(s["SOC"], abs=5e-4)
    r2 = O.simulate_batch(m, th, O.make_run("V", 0.0, input_kind="hold", new_run=False), O.default_opts(), b,
                          state=r["state"])
    s2 = goldens["summaries"]["CV_hold"]
    assert r2["flag"][0] == 4
    assert r2["t_end"][0] == pytest.approx(s2["t_s"], rel=1e-3)
    assert r2["I_end"][0] == pytest.approx(s2["I_C"], abs=2e-3)


def test_jacobian_complex_step_vs_finite_difference():
    m = O.make_model("LCO"); th = O.theta_defaults("LCO"); L = O.layout(m)
    y = O.initial_guess(m, th, 0.5); y[L.I] = 1.0
    it, y, yp = O.newton_init(m, th, O.make_run("I", 1.0), O.default_opts(), y)
    run = O.make_run("I", 1.0)
    cp, rv = O.jac_pattern(m, "I")
    nz = O.jacobian(m, th, run, 0.0, y, yp, 0.7)
    r0 = O.residual(m, th, run, 0.0, y, yp)
    rng = np.random.default_rng(0)
    for c in rng.choice(301, 25, replace=False):
        h = 1e-6 * max(abs(y[c]), 1e-3)
        y2 = y.copy(); y2[c] += h; yp2 = yp.copy(); yp2[c] += 0.7 * h
        fd = (O.residual(m, th, run, 0.0, y2, yp2) - r0) / h
        for k in range(cp[c], cp[c + 1]):
            assert nz[k] == pytest.approx(fd[rv[k]], rel=2e-4, abs=1e-9 * max(1.0, abs(nz[k])))


def test_rng_matches_numpy_copy():
    from tests.util import splitmix_u01
    for sid in (0, 1, 17, 65535):
        for pid in range(7):
            assert O.rng_u01(20211, sid, pid) == splitmix_u01(20211, np.array([sid]), pid)[0]


# ---------------------------------------------------------------------------------------------------
# temperature=true (351 DAEs): examples/fast_charging_CC-CT-CV.ipynb, executed with the current PETLION
# ---------------------------------------------------------------------------------------------------
def _thermal_run():
    m = O.make_model("LCO", temperature=True); th = O.theta_defaults("LCO")
    b = O.default_bounds("LCO", T_max=40 + 273.15, V_max=4.1, I_max=4.0, I_min=1 / 20)
    r = O.simulate_batch(m, th, O.make_run("I", 4.0), O.default_opts(), b, SOC0=0.0, n_save_max=400)
    return m, th, b, r


def test_thermal_layout_and_pattern():
    m = O.make_model("LCO", temperature=True)
    L = O.layout(m)
    assert (L.N_diff, L.N_alg, L.N_tot) == (280, 71, 351)
    assert L.T == 230 and L.j == 280
    cp, rv = O.jac_pattern(m, "I")
    assert len(rv) == 2883


def test_thermal_printed_summary(goldens):
    """simulate(p, I=4) from SOC 0 with T_max = 40 C: every printed digit of the reference's summary."""
    m, th, b, r = _thermal_run()
    s = goldens["summaries"]["thermal_4C_to_Tmax"]
    assert r["flag"][0] == 5                                        # "Above max. temperature"
    assert round(r["t_end"][0], 2) == s["t_s"]
    assert round(r["V_end"][0], 4) == s["V"]
    assert round(r["V_end"][0] * 4.0 * O.calc_I1C(th), 2) == s["P"]
    assert round(r["SOC_end"][0], 4) == s["SOC"]
    assert round(r["T_end"][0] - 273.15, 4) == s["T_C"]


def test_thermal_step_ladder_matches_notebook(goldens):
    """All 76 IDA steps of the thermal 4C charge (SVG polyline of cell 17) are reproduced."""
    m, th, b, r = _thermal_run()
    g = goldens["ladder_4C_thermal"]
    gt, gy = np.array(g["t"]), np.array(g["y_px"])
    n = r["traj_n"][0]
    t, V = r["traj"]["t"][0, :n], r["traj"]["V"][0, :n]
    assert n == len(gt) == 77
    assert np.all(np.abs(t - gt) <= 0.002 + 1e-3 * gt)             # pixel resolution ~1 ms
    # the y pixels are an affine image of V: fit the two calibration constants, check the other 75
    A = np.vstack([gy, np.ones(n)]).T
    c = np.linalg.lstsq(A, V, rcond=None)[0]
    assert np.max(np.abs(A @ c - V)) < 2e-5


def test_thermal_jacobian_complex_step_vs_finite_difference():
    m = O.make_model("LCO", temperature=True); th = O.theta_defaults("LCO"); L = O.layout(m)
    y = O.initial_guess(m, th, 0.5); y[L.I] = 2.0
    y[L.T:L.T + 50] += np.linspace(3.0, 9.0, 50)                    # a temperature profile
    run = O.make_run("I", 2.0)
    it, y, yp = O.newton_init(m, th, run, O.default_opts(), y)
    cp, rv = O.jac_pattern(m, "I")
    nz = O.jacobian(m, th, run, 0.0, y, yp, 0.7)
    r0 = O.residual(m, th, run, 0.0, y, yp)
    for c in list(range(L.T, L.T + 50, 7)) + [0, 9, 29, L.phi_e, L.phi_e + 2, L.phi_s + 2, L.phi_s + 17, L.j + 3, L.I]:
        h = 1e-5 * max(abs(y[c]), 1e-2)                              # central difference
        yb = y.copy(); yb[c] += h; ypb = yp.copy(); ypb[c] += 0.7 * h
        ya = y.copy(); ya[c] -= h; ypa = yp.copy(); ypa[c] -= 0.7 * h
        fd = (O.residual(m, th, run, 0.0, yb, ypb) - O.residual(m, th, run, 0.0, ya, ypa)) / (2 * h)
        for k in range(cp[c], cp[c + 1]):
            assert nz[k] == pytest.approx(fd[rv[k]], rel=1e-4, abs=1e-6 * max(1.0, abs(nz[k]))), (c, rv[k])


# ---------------------------------------------------------------------------------------------------
# dT = :hold (constant-temperature mode) and the CV hold that follows: cells 11 and 13 of the same notebook
# ---------------------------------------------------------------------------------------------------
def test_thermal_CT_hold_start_matches_notebook(goldens):
    """simulate!(sol, p, dT=:hold): the algebraic re-initialisation (I jumps from 4C to 3.28C) and the first
    28 steps reproduce the notebook's current trace to plotting precision."""
    m, th, b, r1 = _thermal_run()
    r2 = O.simulate_batch(m, th, O.make_run("dT", 0.0, input_kind="hold", new_run=False), O.default_opts(), b,
                          state=r1["state"], n_save_max=400)
    g = goldens["trace_CT_hold"]
    gt, gI = np.array(g["t"]), np.array(g["I"])
    t, I = r2["traj"]["t"][0], r2["traj"]["I"][0]
    assert r2["n_newton_init"][0] == 3
    np.testing.assert_allclose(I[:29], gI[:29], atol=3e-4)
    np.testing.assert_allclose(t[:29], gt[:29], atol=2e-3)
    assert abs(I[0] - 3.28) < 5e-4


def test_thermal_CT_CV_printed_summaries(goldens):
    """After step 28 a step-size decision is marginal (rr = 1.998 against the threshold 2: the oracle keeps h,
    the reference doubled it), so from there the two integrations differ at the level of the integration
    tolerance (reltol = 1e-3): printed end values agree to ~1e-3 relative; both bracket the converged
    values (t = 686.69 s, I = 2.7890C; then t = 1875.6 s, I = 0.1910C at reltol 1e-5)."""
    m, th, b, r1 = _thermal_run()
    r2 = O.simulate_batch(m, th, O.make_run("dT", 0.0, input_kind="hold", new_run=False), O.default_opts(), b,
                          state=r1["state"])
    s = goldens["summaries"]["thermal_CT_hold"]
    assert r2["flag"][0] == 2                                        # "Above max. voltage"
    assert r2["t_end"][0] == pytest.approx(s["t_s"], rel=1e-3)
    assert r2["I_end"][0] == pytest.approx(s["I_C"], rel=1e-3)
    assert r2["SOC_end"][0] == pytest.approx(s["SOC"], rel=1e-3)
    assert r2["V_end"][0] == pytest.approx(4.1, abs=1e-9)
    assert r2["T_end"][0] - 273.15 == pytest.approx(s["T_C"], abs=1e-3)   # the mean temperature is held
    r3 = O.simulate_batch(m, th, O.make_run("V", 0.0, input_kind="hold", new_run=False), O.default_opts(), b,
                          state=r2["state"])
    s3 = goldens["summaries"]["thermal_CV_after_CT"]
    assert r3["flag"][0] == 4                                        # "Above max. SOC"
    assert r3["t_end"][0] == pytest.approx(s3["t_s"], rel=1e-2)
    assert r3["I_end"][0] == pytest.approx(s3["I_C"], rel=3e-2)
    assert r3["T_end"][0] - 273.15 == pytest.approx(s3["T_C"], abs=0.05)


def test_dT_jacobian_pattern_and_values():
    m = O.make_model("LCO", temperature=True); th = O.theta_defaults("LCO"); L = O.layout(m)
    cp, rv = O.jac_pattern(m, "dT")
    assert len(rv) == 2883 - 1 + 50                                  # control row: the 50 temperatures
    ctrl_cols = [c for c in range(L.N_tot) if L.I in rv[cp[c]:cp[c + 1]]]
    assert ctrl_cols == list(range(L.T, L.T + 50))
    y = O.initial_guess(m, th, 0.4); y[L.I] = 3.0
    nz = O.jacobian(m, th, O.make_run("dT", 0.0), 0.0, y, np.zeros_like(y), 2.5)
    vals = np.array([nz[cp[c] + list(rv[cp[c]:cp[c + 1]]).index(L.I)] for c in ctrl_cols])
    names = O.theta_names(); d = dict(zip(names, th))
    w = np.repeat([d["l_a"], d["l_p"], d["l_s"], d["l_n"], d["l_z"]], 10) / 10 / (d["l_a"] + d["l_p"] + d["l_s"] + d["l_n"] + d["l_z"])
    np.testing.assert_allclose(vals, -2.5 * w, rtol=1e-12)           # -gamma * temperature_weighting weights


# ---------------------------------------------------------------------------------------------------
# Goldens printed by OLDER PETLION versions (getting_started, CC-CV, variable_input_functions notebooks).
# Those versions did not estimate dY_alg/dt in newtons_method! (Y'_alg = 0 at the start of a run, which
# changes IDA's first step size and hence the whole step ladder).  With that one switch the oracle
# reproduces every printed digit and every decoded step time of those notebooks, which pins the IDA
# restatement (error-test failures, failed-step returns, stop times, re-initialisation) to the reference.
# ---------------------------------------------------------------------------------------------------
@pytest.fixture
def older_version():
    O.lib().orc_debug_skip_alg_deriv(1)
    yield
    O.lib().orc_debug_skip_alg_deriv(0)


def _printed(x, digits):
    return round(float(x), digits)


def test_older_version_1C_discharge_printed_digits(goldens, older_version):
    m = O.make_model("LCO"); th = O.theta_defaults("LCO")
    r = O.simulate_batch(m, th, O.make_run("I", -1.0), O.default_opts(), O.default_bounds("LCO"), SOC0=1.0)
    s = goldens["summaries"]["1C_discharge"]
    assert _printed(r["V_end"][0], 4) == s["V"]                                        # 2.9357
    assert _printed(r["I_end"][0] * O.calc_I1C(th) * r["V_end"][0], 4) == s["P"]       # -85.8094
    assert r["t_end"][0] == pytest.approx(3600.0, abs=1e-6) and r["flag"][0] == 3


def test_older_version_2C_charge_ladder_and_printed_digits(goldens, older_version):
    m = O.make_model("LCO"); th = O.theta_defaults("LCO")
    b = O.default_bounds("LCO", V_max=4.1)
    r = O.simulate_batch(m, th, O.make_run("I", 2.0, tf=1800.0), O.default_opts(), b, SOC0=0.0, n_save_max=400)
    s = goldens["summaries"]["2C_CC_to_4.1V"]
    assert r["flag"][0] == 2
    assert _printed(r["t_end"][0], 2) == s["t_s"]                                      # 1388.68
    assert _printed(r["SOC_end"][0], 4) == s["SOC"]                                    # 0.7715
    assert _printed(r["I_end"][0] * O.calc_I1C(th) * r["V_end"][0], 4) == s["P"]       # 239.6861
    gt = np.array(goldens["ladder_CCCV_older_version"]["t"][0])
    n = r["traj_n"][0]
    assert n == 84                                       # CC-CV.ipynb cell 9: 84 points, then the CV hold
    t = r["traj"]["t"][0, :n]
    assert np.all(np.abs(t - gt[:n]) <= 0.02 + 2e-3 * gt[:n])      # 500 s ticks: ~10 ms per pixel-thousandth


@pytest.mark.parametrize("name", ["step", "step_tdiscon", "ramp_100", "ramp_10"])
def test_older_version_function_inputs_printed_digits(goldens, older_version, name):
    """examples/variable_input_functions.ipynb: a discontinuous current with and without tdiscon (the run
    crosses the jump through error-test failures, failed IDA returns and check_reinitialization!), and two
    current ramps with their full step ladders"""
    g = goldens["function_inputs"][name]
    m = O.make_model("LCO"); th = O.theta_defaults("LCO")
    if name.startswith("step"):
        run = O.make_run("I", tf=200, table=([0.0, 100.0, 100.0], [1.0, 1.0, 0.5]), tdiscon=g["tdiscon"])
    else:
        run = O.make_run("I", tf=100, table=([0.0, 100.0], [0.0, 100.0 * g["ramp_val"]]))
    r = O.simulate_batch(m, th, run, O.default_opts(), O.default_bounds("LCO"), SOC0=0.0, n_save_max=400)
    assert r["flag"][0] == 0 and r["t_end"][0] == g["t_s"]
    assert _printed(r["V_end"][0], 4) == g["V"]
    assert _printed(r["I_end"][0] * O.calc_I1C(th) * r["V_end"][0], 4) == g["P"]
    assert _printed(r["SOC_end"][0], 4) == g["SOC"]
    if name == "step":
        assert r["n_reinit"][0] == 1
    if name == "step_tdiscon":
        n = r["traj_n"][0]
        assert np.any(np.abs(r["traj"]["t"][0, :n] - (100.0 - 0.5e-3)) < 1e-9)        # the stop at tdiscon - reltol/2
    if "t_ladder" in g:
        gt = np.array(g["t_ladder"]); n = r["traj_n"][0]
        assert n == len(gt)                                                            # 30 / 58 points
        assert np.all(np.abs(r["traj"]["t"][0, :n] - gt) <= 2e-3 + 1e-3 * gt)


def test_function_input_table_semantics():
    tab = ([0.0, 100.0, 100.0, 200.0], [1.0, 1.0, 0.5, 0.25])
    assert O.table_eval(tab, -5.0) == 1.0 and O.table_eval(tab, 99.999999) == 1.0
    assert O.table_eval(tab, 100.0) == 0.5                # right-continuous at the jump: t < 100 ? 1 : 0.5
    assert O.table_eval(tab, 150.0) == pytest.approx(0.375) and O.table_eval(tab, 1e9) == 0.25
    # a constant table is the same run as a number (test/runtests.jl:35)
    m = O.make_model("LCO"); th = O.theta_defaults("LCO")
    a = O.simulate_batch(m, th, O.make_run("I", 1.0, tf=500), O.default_opts(), O.default_bounds("LCO"), SOC0=0.0)
    b = O.simulate_batch(m, th, O.make_run("I", tf=500, table=([0.0, 1e3], [1.0, 1.0])), O.default_opts(),
                         O.default_bounds("LCO"), SOC0=0.0)
    assert a["n_steps"][0] == b["n_steps"][0] and a["V_end"][0] == b["V_end"][0]


def test_function_input_current_version(goldens):
    """the same notebook cases with the current reference semantics (Y'_alg estimated): 4-digit agreement"""
    m = O.make_model("LCO"); th = O.theta_defaults("LCO")
    for name, tol in (("step", 1.5e-4), ("step_tdiscon", 1.5e-4), ("ramp_100", 1.5e-4), ("ramp_10", 1.5e-4)):
        g = goldens["function_inputs"][name]
        if name.startswith("step"):
            run = O.make_run("I", tf=200, table=([0.0, 100.0, 100.0], [1.0, 1.0, 0.5]), tdiscon=g["tdiscon"])
        else:
            run = O.make_run("I", tf=100, table=([0.0, 100.0], [0.0, 100.0 * g["ramp_val"]]))
        r = O.simulate_batch(m, th, run, O.default_opts(), O.default_bounds("LCO"), SOC0=0.0)
        assert abs(r["V_end"][0] - g["V"]) < tol and abs(r["SOC_end"][0] - g["SOC"]) < 5e-5


def test_user_stop_times_merge_like_postfix_integrator():
    """tstops = sort([opts.tstops; tdiscon .- reltol/2; 1.0 if continuing; tf]) (model_evaluation.jl:288-310)"""
    m = O.make_model("LCO"); th = O.theta_defaults("LCO")
    r = O.simulate_batch(m, th, O.make_run("I", -1.0, tf=1000, tstops=[250.5, 100.0, 2000.0, -3.0]), O.default_opts(),
                         O.default_bounds("LCO"), SOC0=1.0, n_save_max=400)
    t = r["traj"]["t"][0, :r["traj_n"][0]]
    assert np.any(t == 100.0) and np.any(t == 250.5) and t[-1] == 1000.0 and r["flag"][0] == 0
    r2 = O.simulate_batch(m, th, O.make_run("I", 1.0, tf=300, tstops=[0.25, 40.0], new_run=False), O.default_opts(),
                          O.default_bounds("LCO"), state=r["state"], n_save_max=400)
    t2 = r2["traj"]["t"][0, :r2["traj_n"][0]] - 1000.0
    assert all(np.any(np.abs(t2 - x) < 1e-9) for x in (0.25, 1.0, 40.0))


def test_exit_blend_is_second_order():
    """The END of a run that trips a bound is the linear blend of the last step (interp_final_points!,
    model_evaluation.jl:369-382): h^2-accurate in the last step size whatever the tolerance.  The oracle against
    itself at 1e-9 and 1e-10: the trajectory (dense rows, both sides' own interpolant) agrees to 1e-8, the end values
    only to ~1e-5 -- which is why tests/test_gpu_tight.py holds end values of runs with different last steps to
    BLEND_TOL instead of 1e-6."""
    m = O.make_model("LCO")
    th = O.theta_defaults("LCO")[None, :]
    td = np.arange(0.0, 3590.0, 100.0)
    r = {}
    for tol in (1e-9, 1e-10):
        o = O.default_opts(reltol=tol, abstol=tol, reltol_init=tol, abstol_init=tol)
        r[tol] = O.simulate_batch(m, th, O.make_run("I", -2.0), o, O.default_bounds("LCO"), SOC0=1.0, dense_t=td)
    a, b = r[1e-9], r[1e-10]
    assert a["flag"][0] == b["flag"][0] == 1 and a["n_steps"][0] != b["n_steps"][0]          # 2C: ends on V_min
    fill = ~np.isnan(a["dense"]["V"]) & ~np.isnan(b["dense"]["V"])
    assert np.max(np.abs(a["dense"]["V"][fill] - b["dense"]["V"][fill]) / b["dense"]["V"][fill]) < 2e-8
    dt = abs(a["t_end"][0] - b["t_end"][0]) / b["t_end"][0]
    assert 1e-8 < dt < 2e-4


def test_norm_summation_order_is_not_what_flips_decisions():
    """SURVEY.md 7 H1 / VERDICT r1 next-1c: would the oracle agree with the kernel on more step sequences if its WRMS
    norms were summed in the kernel's order (32 lane partials with fma, xor-butterfly tree)?  No: switching the oracle
    itself between the serial sum and that order changes the last bits of most results but the DECISIONS (every
    counter of the run) of about one system in four thousand -- a hundred times less than the ~3 % of systems on which
    GPU and oracle part ways.  Those come from the different (both exact) linear solvers acting on the Newton
    corrections, not from the norms."""
    import ctypes as C
    m = O.make_model("LCO")
    B = 1024
    from tests import util
    tho = util.oracle_theta_batch(B, first=3000)
    L = O.lib()
    run = lambda: O.simulate_batch(m, tho, O.make_run("I", -1.0), O.default_opts(), O.default_bounds("LCO"), SOC0=1.0, nthreads=8)
    r0 = run()
    L.orc_debug_norm_mode(1)
    try:
        r1 = run()
    finally:
        L.orc_debug_norm_mode(0)
    same = np.all([r0[k] == r1[k] for k in ("n_steps", "flag", "n_res", "n_jac", "n_netf", "n_ncfn")], axis=0)
    assert same.mean() >= 0.995                       # observed: 4095 of 4096
    assert np.mean(r0["V_end"] == r1["V_end"]) < 0.9   # ... while the bits do change
    ok = r0["flag"] >= 0
    np.testing.assert_allclose(r1["V_end"][ok & same], r0["V_end"][ok & same], rtol=1e-7)


def test_tight_tolerance_cv_phase_is_erratic_in_the_oracle_itself():
    """Why tests/test_gpu_tight.py cannot hold the thermal CV phase to 1e-6 for 100 % of the systems: the restated IDA
    (Sundials.jl's settings: 3 Newton iterations, acceptance on the last rate estimate) is not monotone in the tolerance
    there.  4C charge to 4.1 V then V = :hold on that test's batch: at reltol 1e-7 the CV current of about one system
    in twenty sits 1-9 % away from its value at 1e-8 (and at 1e-6, 1e-9), the rest agree to 1e-4.  WHICH systems is
    round-off luck -- it changes with the compiler's contraction choices: one oracle build had system 237 off by 9 %
    (the GPU agreed with the 1e-8 value there), the next build 26 other systems of the 512 -- so the test counts them
    instead of naming one."""
    from tests import util
    W = util.PROTOCOLS["cfg3i"]
    tho = util.oracle_theta_batch(128, first=80000)
    td = np.array([1500.0, 1700.0, 1845.0, 2000.0])
    cur = {}
    for tol in (1e-7, 1e-8):
        o = O.default_opts(reltol=tol, abstol=tol, reltol_init=tol, abstol_init=tol, maxiters=400000)
        cur[tol] = util.oracle_protocol(W, tho, o, dense_t=td, nthreads=16)[1]["dense"]["I"]
    d = np.abs(cur[1e-7] - cur[1e-8]) / np.abs(cur[1e-8])
    d = np.where(np.isnan(d), 0.0, d).max(axis=1)
    assert np.mean(d <= 1e-4) >= 0.8          # most systems: the two tolerances agree
    assert (d > 1e-2).sum() >= 1              # ... but not all, and not by a little
    assert d.max() < 0.2                      # (the loose bound tests/test_gpu_tight.py uses for the CV rows)
