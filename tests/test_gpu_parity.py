"""GPU parity tests (through the C ABI): CUDA path vs the CPU oracle on the same seeded inputs.

Tolerances (FP64 path, north_star: rtol 1e-6 on voltage/SOC trajectories):
  residual / Jacobian values : 1e-10 relative to the row/entry scale (round-off only)
  Newton initialisation      : 1e-9
  trajectories               : identical step counts; V, SOC within rtol 1e-6 (observed ~1e-9)
"""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu

# (offset by 7 s: most runs end at t = 3600 s to round-off, a requested time there is filled by one side only)
DENSE_T = np.concatenate([[0.0], np.arange(7.0, 3700.0, 30.0), [1e6]])


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


@pytest.fixture(scope="module")
def lco(P):
    return P.petlion("LCO")


def test_theta_keys_reference_order(lco):
    # get_symbolic_vars sorts the keys (generate_functions.jl:387); Julia sorts Symbols by code point
    assert lco.θ_keys == sorted(lco.θ_keys)
    assert len(lco.θ_keys) == 35
    od = O.theta_dict("LCO")
    th = util.product_theta_from_oracle(lco, np.array([list(od.values())]))
    assert np.array_equal(th[0], np.array(list(lco.θ.values())))


@pytest.mark.parametrize("method", ["I", "V", "P"])
def test_jac_pattern_equals_oracle(lco, method):
    m = O.make_model("LCO")
    cp, rv = O.jac_pattern(m, method)
    cp2, rv2 = lco.jac_pattern(method)
    assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)


def test_initial_guess(lco):
    m = O.make_model("LCO")
    tho = util.oracle_theta_batch(16)
    th = util.product_theta_from_oracle(lco, tho)
    soc = np.linspace(0, 1, 16)
    Y0 = lco.initial_guess(soc, theta=th)
    for s in range(16):
        ref = O.initial_guess(m, tho[s], soc[s])
        np.testing.assert_allclose(Y0[s], ref, rtol=1e-13, atol=0)


@pytest.mark.parametrize("method,value", [("I", 1.0), ("V", 3.9), ("P", 50.0)])
def test_resjac_parity(lco, method, value):
    m = O.make_model("LCO")
    B = 48
    tho = util.oracle_theta_batch(B)
    # include a non-reference temperature so the Arrhenius / dU/dT branches are exercised
    tho[B // 2:, O.theta_names().index("T0")] = 305.0
    th = util.product_theta_from_oracle(lco, tho)
    Y, YP = util.random_states(m, tho, seed=3)
    gam = np.random.default_rng(5).uniform(0.01, 50.0, size=B)
    res, nz = lco.resjac(Y, YP, gam, method=method, value=value, theta=th)
    run = O.make_run(method, value)
    cp, rv = O.jac_pattern(m, method)
    for s in range(B):
        r_ref = O.residual(m, tho[s], run, 0.0, Y[s], YP[s])
        j_ref = O.jacobian(m, tho[s], run, 0.0, Y[s], YP[s], gam[s])
        # scale of a residual row: the largest |J_ij * y_j| contribution in that row
        scale = np.zeros(301)
        for c in range(301):
            k = slice(cp[c], cp[c + 1])
            np.maximum.at(scale, rv[k], np.abs(j_ref[k]) * max(abs(Y[s][c]), 1e-12))
        scale = np.maximum(scale, np.abs(r_ref))
        assert np.all(np.abs(res[s] - r_ref) <= 1e-10 * scale + 1e-300), (s, np.argmax(np.abs(res[s] - r_ref) / scale))
        # Jacobian entries: relative to the largest entry of the same row
        rowmax = np.zeros(301)
        np.maximum.at(rowmax, rv, np.abs(j_ref))
        err = np.abs(nz[s] - j_ref) / rowmax[rv]
        assert err.max() < 1e-9, (s, int(np.argmax(err)), err.max())


def test_newton_init_parity(lco):
    m = O.make_model("LCO"); L = O.layout(m)
    B = 32
    tho = util.oracle_theta_batch(B)
    th = util.product_theta_from_oracle(lco, tho)
    soc = np.linspace(0.05, 0.95, B)
    cur = np.where(np.arange(B) % 2 == 0, -1.0, 2.0)
    Y0 = lco.initial_guess(soc, theta=th)
    Y0[:, L.I] = cur
    st, Y, YP = lco.newton_init(Y0, method="I", value=cur, theta=th)
    opts = O.default_opts()
    for s in range(B):
        it, y, yp = O.newton_init(m, tho[s], O.make_run("I", cur[s]), opts, Y0[s])
        assert st[s] == it
        np.testing.assert_allclose(Y[s], y, rtol=1e-9, atol=1e-12)
        scale = np.maximum(np.abs(yp), 1e-6 * np.abs(yp).max())
        assert np.max(np.abs(YP[s] - yp) / scale) < 1e-6


def _compare_runs(sol, ref, rtol=1e-6, min_identical=1.0, flip_tol=5e-3):
    summ = sol.results[-1].summary
    same_steps = summ["n_steps"] == ref["n_steps"]
    assert np.mean(same_steps) >= min_identical, (np.mean(same_steps), np.where(~same_steps)[0][:10])
    assert np.array_equal(summ["flag"][same_steps], ref["flag"][same_steps])
    idx = np.where(same_steps)[0]
    np.testing.assert_allclose(summ["t_end"][idx], ref["t_end"][idx], rtol=rtol)
    np.testing.assert_allclose(summ["V_end"][idx], ref["V_end"][idx], rtol=rtol)
    np.testing.assert_allclose(summ["SOC_end"][idx], ref["SOC_end"][idx], rtol=rtol, atol=1e-8)
    worst = 0.0
    for s in idx:
        n = ref["traj_n"][s]
        assert sol.n_points[s] >= n or True
        tv = sol.V[s, :n]; rv = ref["traj"]["V"][s, :n]
        worst = max(worst, np.max(np.abs(tv - rv) / np.abs(rv)))
        np.testing.assert_allclose(sol.t[s, :n], ref["traj"]["t"][s, :n], rtol=rtol, atol=1e-9)
        np.testing.assert_allclose(sol.SOC[s, :n], ref["traj"]["SOC"][s, :n], rtol=rtol, atol=1e-8)
    assert worst < rtol, worst
    if sol.dense is not None and ref.get("dense") is not None:
        bound_flipped(sol.results[-1].summary, sol.dense, ref, tol=flip_tol)
    else:
        bound_flipped_rows(sol, ref, tol=flip_tol)
    return worst


def bound_flipped_rows(sol, ref, tol=5e-3, margin=30.0):
    """the same bound from the saved rows when no dense output was requested: the GPU's rows of the LAST run of `sol`
    interpolated linearly at the oracle's step times"""
    summ = sol.results[-1].summary
    assert np.array_equal(summ["flag"] < 0, ref["flag"] < 0)
    okb = (summ["flag"] >= 0) & (ref["flag"] >= 0)
    n_new = sol.results[-1].n_rows
    for i in np.where(okb & ((summ["n_steps"] != ref["n_steps"]) | (summ["flag"] != ref["flag"])))[0]:
        a0 = sol.n_points[i] - n_new[i]
        gt, gV, gS = (x[i, a0:sol.n_points[i]] for x in (sol.t, sol.V, sol.SOC))
        n = ref["traj_n"][i]
        rt, rV, rS = ref["traj"]["t"][i, :n], ref["traj"]["V"][i, :n], ref["traj"]["SOC"][i, :n]
        sel = rt <= min(gt[-1], rt[-1]) - margin
        if not sel.any():
            continue
        assert np.max(np.abs(np.interp(rt[sel], gt, gV) - rV[sel]) / np.abs(rV[sel])) <= tol
        assert np.max(np.abs(np.interp(rt[sel], gt, gS) - rS[sel])) <= tol
        assert abs(summ["t_end"][i] - ref["t_end"][i]) <= tol * max(ref["t_end"][i] - rt[0], 1.0)


def bound_flipped(summ, dense, ref, tol=5e-3, margin=30.0):
    """The systems that took a different discrete decision somewhere (a Newton-convergence or error-test threshold
    crossed by round-off) are two runs of an adaptive integrator at tolerance reltol = 1e-3: they must still agree to
    a few reltol (tol = 5 reltol).  Compared on a common grid through the dense output of both sides, up to `margin` seconds before the
    earlier of the two ends (the last step of a run is the linear blend of interp_final_points!, and where one side
    exits on V_min and the other on SOC_min the voltage falls by volts per minute).  Returns the worst values."""
    assert dense is not None and ref.get("dense") is not None, "run both sides with dense output"
    okb = (summ["flag"] >= 0) & (ref["flag"] >= 0)
    flipped = okb & ~np.all([summ[k] == ref[k] for k in ("n_steps", "flag", "n_res", "n_jac", "n_netf", "n_ncfn")], axis=0)
    # hard failures: both sides must agree on WHICH systems fail
    assert np.array_equal(summ["flag"] < 0, ref["flag"] < 0), (np.where((summ["flag"] < 0) != (ref["flag"] < 0))[0][:10])
    td = dense["t"]
    worst = dict(n_flipped=int(flipped.sum()), dV=0.0, dSOC=0.0, dt_end=0.0)
    for i in np.where(flipped)[0]:
        sel = td <= min(summ["t_end"][i], ref["t_end"][i]) - margin
        a, b = dense["V"][i, sel], ref["dense"]["V"][i, sel]
        assert not np.isnan(a).any() and not np.isnan(b).any()
        worst["dV"] = max(worst["dV"], float(np.max(np.abs(a - b) / np.abs(b))))
        worst["dSOC"] = max(worst["dSOC"], float(np.max(np.abs(dense["SOC"][i, sel] - ref["dense"]["SOC"][i, sel]))))
        worst["dt_end"] = max(worst["dt_end"], abs(summ["t_end"][i] - ref["t_end"][i]) / ref["t_end"][i])
    assert worst["dV"] <= tol and worst["dSOC"] <= tol and worst["dt_end"] <= tol, worst
    return worst


def test_simulate_nominal_1C_discharge(P, lco, goldens):
    """configs[0]: single LCO 1C CC discharge, SOC 1 -> 0: 80 steps, exit on SOC_min at t = 3600 s."""
    m = O.make_model("LCO")
    for k, v in zip(lco.θ_keys, P.petlion("LCO").θ.values()):
        lco.θ[k] = v
    sol = P.simulate(lco, I=-1, SOC=1, dense_t=DENSE_T)
    ref = O.simulate_batch(m, O.theta_defaults("LCO"), O.make_run("I", -1.0), O.default_opts(),
                           O.default_bounds("LCO"), SOC0=1.0, n_save_max=512, dense_t=DENSE_T)
    s = sol.results[-1].summary
    assert s["n_steps"][0] == 80 and s["flag"][0] == 3
    assert abs(s["t_end"][0] - 3600.0) < 1e-6
    worst = _compare_runs(sol, ref)
    # and against the notebook ladder itself
    g = goldens["ladder_1C_discharge"]["eps_p"]["0.385"]
    gt = np.array(g["t"])
    assert np.all(np.abs(sol.t[0, :81] - gt) <= 0.004 + 2e-3 * gt)


def test_simulate_randomised_batch_parity(P, lco):
    """configs[1] at a size the oracle finishes in seconds: randomised {D_s, k, eps} batch."""
    m = O.make_model("LCO")
    B = 192
    tho = util.oracle_theta_batch(B)
    th = util.product_theta_from_oracle(lco, tho)
    util.set_theta_batch(lco, th)
    sol = P.simulate(lco, I=-1, SOC=1, dense_t=DENSE_T)
    ref = O.simulate_batch(m, tho, O.make_run("I", -1.0), O.default_opts(), O.default_bounds("LCO"),
                           SOC0=1.0, n_save_max=512, nthreads=8, dense_t=DENSE_T)
    assert np.all(np.isin(ref["flag"], (1, 3)))      # SOC_min, or V_min for the slow-diffusion draws
    _compare_runs(sol, ref, min_identical=0.98)


def test_flipped_systems_are_bounded_on_a_common_grid(P, lco):
    """4 096 randomised systems: EVERY system is accounted for -- identical step sequence (V to 1e-6), or a different
    sequence whose V(t), SOC(t) agree to 5*reltol on a common grid, or a hard failure that the oracle shares."""
    m = O.make_model("LCO")
    B = 4096
    tho = util.oracle_theta_batch(B, first=20000)
    util.set_theta_batch(lco, util.product_theta_from_oracle(lco, tho))
    sol = P.simulate(lco, I=-1, SOC=1, dense_t=DENSE_T, n_save_max=0)
    ref = O.simulate_batch(m, tho, O.make_run("I", -1.0), O.default_opts(), O.default_bounds("LCO"),
                           SOC0=1.0, nthreads=16, dense_t=DENSE_T)
    s = sol.results[-1].summary
    # identical sequence: every counter of the run agrees (steps, residual and Jacobian evaluations, failures)
    same = (ref["flag"] >= 0) & np.all([s[k] == ref[k] for k in ("n_steps", "flag", "n_res", "n_jac", "n_netf", "n_ncfn")], axis=0)
    assert same.mean() > 0.9
    fill = ~np.isnan(ref["dense"]["V"][same]) & ~np.isnan(sol.dense["V"][same])
    assert (np.isnan(ref["dense"]["V"][same]) != np.isnan(sol.dense["V"][same])).sum(axis=1).max() <= 1
    np.testing.assert_allclose(sol.dense["V"][same][fill], ref["dense"]["V"][same][fill], rtol=1e-6)
    np.testing.assert_allclose(s["t_end"][same], ref["t_end"][same], rtol=1e-6)
    w = bound_flipped(s, sol.dense, ref)
    print("identical", same.mean(), "flipped", w, "hard failures", int((ref["flag"] < 0).sum()))


def test_simulate_cccv_continuation(P, lco):
    """2C CC charge to 4.1 V then V=:hold until SOC_max (simulate! continuation, examples/CC-CV.ipynb)."""
    m = O.make_model("LCO")
    B = 24
    tho = util.oracle_theta_batch(B, first=1000)
    th = util.product_theta_from_oracle(lco, tho)
    util.set_theta_batch(lco, th)
    sol = P.simulate(lco, I=2, SOC=0, V_max=4.1, dense_t=DENSE_T)
    b = O.default_bounds("LCO", V_max=4.1)
    ref = O.simulate_batch(m, tho, O.make_run("I", 2.0), O.default_opts(), b, SOC0=0.0, n_save_max=512, nthreads=8,
                           dense_t=DENSE_T)
    assert np.all(ref["flag"] == 2)
    _compare_runs(sol, ref, min_identical=0.9)
    n1 = sol.n_points.copy()
    P.simulate_(sol, lco, V="hold", V_max=4.1)
    ref2 = O.simulate_batch(m, tho, O.make_run("V", 0.0, input_kind="hold", new_run=False), O.default_opts(), b,
                            state=ref["state"], n_save_max=512, nthreads=8)
    s2 = sol.results[-1].summary
    same = s2["n_steps"] == ref2["n_steps"]
    assert np.mean(same) >= 0.8
    assert np.array_equal(s2["flag"][same], ref2["flag"][same])
    np.testing.assert_allclose(s2["t_end"][same], ref2["t_end"][same], rtol=1e-6)
    np.testing.assert_allclose(s2["I_end"][same], ref2["I_end"][same], rtol=1e-5, atol=1e-9)
    assert np.all(sol.n_points == n1 + s2["n_steps"] + 1)


def test_tight_tolerance_agreement(P, lco):
    """H1: at reltol=abstol=1e-9 any two correct integrators agree to 1e-6 regardless of step choices."""
    m = O.make_model("LCO")
    B = 8
    tho = util.oracle_theta_batch(B, first=5000)
    th = util.product_theta_from_oracle(lco, tho)
    util.set_theta_batch(lco, th)
    sol = P.simulate(lco, 1800.0, I=-1, SOC=1, abstol=1e-9, reltol=1e-9, n_save_max=4096)
    ref = O.simulate_batch(m, tho, O.make_run("I", -1.0, tf=1800.0), O.default_opts(abstol=1e-9, reltol=1e-9,
                           abstol_init=1e-9, reltol_init=1e-9), O.default_bounds("LCO"), SOC0=1.0, nthreads=8)
    s = sol.results[-1].summary
    np.testing.assert_allclose(s["V_end"], ref["V_end"], rtol=1e-6)
    np.testing.assert_allclose(s["t_end"], ref["t_end"], rtol=1e-9)


def test_fault_isolation(P, lco):
    """poisoning theta of a few systems must not change the others (bit-exact) and must flag the bad ones"""
    B = 64
    tho = util.oracle_theta_batch(B, first=9000)
    th = util.product_theta_from_oracle(lco, tho)
    util.set_theta_batch(lco, th)
    good = P.simulate(lco, I=-1, SOC=1)
    th2 = th.copy()
    bad = [3, 17, 40]
    th2[bad, lco.θ_keys.index("D_sp")] = np.nan
    util.set_theta_batch(lco, th2)
    mixed = P.simulate(lco, I=-1, SOC=1)
    sg, sm = good.results[-1].summary, mixed.results[-1].summary
    ok = np.setdiff1d(np.arange(B), bad)
    assert np.array_equal(sg[ok], sm[ok])
    assert np.array_equal(good.V[ok], mixed.V[ok], equal_nan=True)
    assert np.all(sm["flag"][bad] < 0)


def test_nmc_variant(P):
    m = O.make_model("NMC")
    p = P.petlion("NMC")
    assert len(p.θ_keys) == 32
    od = O.theta_dict("NMC")
    tho = np.array([list(od.values())])
    th = util.product_theta_from_oracle(p, tho)
    Y, YP = util.random_states(m, tho, seed=11)
    res, nz = p.resjac(Y, YP, 0.3, method="I", value=1.0, theta=th)
    run = O.make_run("I", 1.0)
    r_ref = O.residual(m, tho[0], run, 0.0, Y[0], YP[0])
    j_ref = O.jacobian(m, tho[0], run, 0.0, Y[0], YP[0], 0.3)
    cp, rv = O.jac_pattern(m, "I")
    cp2, rv2 = p.jac_pattern("I")
    assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
    rowmax = np.zeros(301); np.maximum.at(rowmax, rv, np.abs(j_ref))
    assert np.max(np.abs(nz[0] - j_ref) / rowmax[rv]) < 1e-9
    scale = np.maximum(np.abs(r_ref), rowmax * np.abs(Y[0]).max() * 1e-6)
    assert np.all(np.abs(res[0] - r_ref) <= 1e-9 * np.maximum(scale, 1e-30))
    # GITT-like pulse + rest (examples/GITT.ipynb): 1C for 180 s then I = :rest for 600 s
    sol = P.simulate(p, 180.0, I=1, SOC=0)
    ref = O.simulate_batch(m, tho, O.make_run("I", 1.0, tf=180.0), O.default_opts(), O.default_bounds("NMC"), SOC0=0.0, n_save_max=512)
    _compare_runs(sol, ref)
    P.simulate_(sol, p, 600.0, I="rest")
    ref2 = O.simulate_batch(m, tho, O.make_run("I", 0.0, tf=600.0, input_kind="rest", new_run=False), O.default_opts(),
                            O.default_bounds("NMC"), state=ref["state"], n_save_max=512)
    s2 = sol.results[-1].summary
    assert s2["flag"][0] == ref2["flag"][0] == 0
    assert s2["n_steps"][0] == ref2["n_steps"][0]
    np.testing.assert_allclose(s2["V_end"], ref2["V_end"], rtol=1e-6)


def test_simulate_power_control(P, lco):
    """method_P (constant power, scalar_residual.jl:189-197): the control row has a zero... no, a V*I1C
    diagonal and two Phi_s entries; exercised through the Schur-complement border."""
    m = O.make_model("LCO")
    B = 16
    tho = util.oracle_theta_batch(B, first=300)
    th = util.product_theta_from_oracle(lco, tho)
    util.set_theta_batch(lco, th)
    sol = P.simulate(lco, 900.0, P=-60.0, SOC=0.9)
    ref = O.simulate_batch(m, tho, O.make_run("P", -60.0, tf=900.0), O.default_opts(), O.default_bounds("LCO"),
                           SOC0=0.9, n_save_max=512, nthreads=8)
    assert np.all(ref["flag"] == 0)
    _compare_runs(sol, ref, min_identical=0.9)
    # power is held: P = I * I1C * V
    I1C = lco.I1C(B)
    s = sol.results[-1].summary
    np.testing.assert_allclose(s["I_end"] * I1C * s["V_end"], -60.0, rtol=1e-6)


def test_gitt_protocol_nmc(P):
    """configs[3] protocol at test size: NMC, SOC0 = 0, repeated {I=+1 for 180 s ; I=:rest for 1200 s}
    through simulate!/continuation (examples/GITT.ipynb)."""
    m = O.make_model("NMC")
    p = P.petlion("NMC")
    B = 6
    names = O.theta_names()
    tho = np.tile(O.theta_defaults("NMC"), (B, 1))
    u = util.splitmix_u01(util.SEED, np.arange(B), 3)
    tho[:, names.index("D_sp")] *= 10.0 ** (0.3 * (2 * u - 1))
    tho[:, names.index("k_n")] *= 10.0 ** (0.3 * (2 * util.splitmix_u01(util.SEED, np.arange(B), 4) - 1))
    th = util.product_theta_from_oracle(p, tho)
    util.set_theta_batch(p, th)
    opts, bounds = O.default_opts(), O.default_bounds("NMC")
    sol, state = None, None
    for cyc in range(3):
        sol = P.simulate(p, 180.0, I=1, SOC=0) if sol is None else P.simulate_(sol, p, 180.0, I=1)
        ref = O.simulate_batch(m, tho, O.make_run("I", 1.0, tf=180.0, new_run=state is None), opts, bounds,
                               SOC0=0.0, state=state, n_save_max=512)
        state = ref["state"]
        s = sol.results[-1].summary
        assert np.array_equal(s["flag"], ref["flag"]) and np.all(s["flag"] == 0)
        assert np.mean(s["n_steps"] == ref["n_steps"]) >= 0.8
        np.testing.assert_allclose(s["V_end"], ref["V_end"], rtol=1e-6)
        np.testing.assert_allclose(s["SOC_end"], ref["SOC_end"], rtol=1e-6, atol=1e-9)
        P.simulate_(sol, p, 1200.0, I="rest")
        ref = O.simulate_batch(m, tho, O.make_run("I", 0.0, tf=1200.0, input_kind="rest", new_run=False), opts, bounds,
                               state=state, n_save_max=512)
        state = ref["state"]
        s = sol.results[-1].summary
        assert np.array_equal(s["flag"], ref["flag"]) and np.all(s["flag"] == 0)
        np.testing.assert_allclose(s["V_end"], ref["V_end"], rtol=1e-6)
        np.testing.assert_allclose(s["t_end"], ref["t_end"], rtol=1e-9)
    assert len(sol.results) == 6
    # global time is continuous across runs
    assert np.all(np.diff(sol.t[0, :sol.n_points[0]]) >= 0)


def test_eta_p_control(P, lco):
    """method_eta_p (scalar_residual.jl:92, 199-203; input_methods.jl:108-143): hold the plating overpotential
    Phi_s.n[1] - Phi_e.n[1]; control row +1/-1 on two interior unknowns, zero corner, through the border."""
    m = O.make_model("LCO"); L = O.layout(m)
    cp, rv = O.jac_pattern(m, "η_p")
    cp2, rv2 = lco.jac_pattern("η_p")
    assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
    B = 12
    tho = util.oracle_theta_batch(B, first=700)
    th = util.product_theta_from_oracle(lco, tho)
    util.set_theta_batch(lco, th)
    b = O.default_bounds("LCO", V_max=4.2)
    # 2C charge for 600 s, then hold the overpotential reached (eta_p = :hold) for another 600 s
    sol = P.simulate(lco, 600.0, I=2, SOC=0.2, V_max=4.2)
    r1 = O.simulate_batch(m, tho, O.make_run("I", 2.0, tf=600.0), O.default_opts(), b, SOC0=0.2, n_save_max=512, nthreads=8)
    _compare_runs(sol, r1, min_identical=0.9)
    P.simulate_(sol, lco, 600.0, eta_p="hold", V_max=4.2)
    r2 = O.simulate_batch(m, tho, O.make_run("η_p", 0.0, tf=600.0, input_kind="hold", new_run=False), O.default_opts(), b,
                          state=r1["state"], n_save_max=512, nthreads=8)
    s2 = sol.results[-1].summary
    same = (s2["n_steps"] == r2["n_steps"]) & (s2["flag"] == r2["flag"])
    assert np.mean(same) >= 0.8, (s2["n_steps"], r2["n_steps"])
    np.testing.assert_allclose(s2["V_end"][same], r2["V_end"][same], rtol=1e-6)
    np.testing.assert_allclose(s2["I_end"][same], r2["I_end"][same], rtol=1e-6)
    eta1 = r1["state"]["Y"][:, L.phi_s + 10] - r1["state"]["Y"][:, L.phi_e + 20]
    eta2 = sol.Y[:, L.phi_s + 10] - sol.Y[:, L.phi_e + 20]
    np.testing.assert_allclose(eta2, eta1, rtol=1e-6)                  # held
    # operator level: residual + Jacobian with the eta_p row, and a fresh run with a number
    Y, YP = r1["state"]["Y"], r1["state"]["YP"]
    res, nz = lco.resjac(Y, YP, 0.3, method="η_p", value=0.05, theta=th)
    for s in range(B):
        run = O.make_run("η_p", 0.05)
        np.testing.assert_allclose(res[s][L.I], O.residual(m, tho[s], run, 0.0, Y[s], YP[s])[L.I], rtol=1e-12)
        j_ref = O.jacobian(m, tho[s], run, 0.0, Y[s], YP[s], 0.3)
        rowmax = np.zeros(301); np.maximum.at(rowmax, rv, np.abs(j_ref))
        assert np.max(np.abs(nz[s] - j_ref) / rowmax[rv]) < 1e-9
    sol3 = P.simulate(lco, 300.0, eta_p=0.08, SOC=0.3, V_max=4.2)
    r3 = O.simulate_batch(m, tho, O.make_run("η_p", 0.08, tf=300.0), O.default_opts(), b, SOC0=0.3, n_save_max=512, nthreads=8)
    s3 = sol3.results[-1].summary
    same = (s3["n_steps"] == r3["n_steps"]) & (s3["flag"] == r3["flag"])
    assert np.mean(same) >= 0.8
    np.testing.assert_allclose(s3["I_end"][same], r3["I_end"][same], rtol=1e-6)
