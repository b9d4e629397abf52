"""Fickian_method = :spectral (params.jl:142; residuals_c_s_avg!, residuals.jl:181-235, labelled "BETA" there): Chebyshev
collocation in the particles -- a dense constant particle block and a surface-flux coupling on EVERY radial row.  Sibling
builds of the isothermal, thermal and SEI families (N_r = 10).  Pattern, residual/Jacobian values, the structured linear
solve (against dense LAPACK), the Newton initialisation and a CC charge + CV hold against the oracle
(tests/test_oracle_spectral.py checks the oracle's restatement against the closed form the scheme implies; the reference
executes the option nowhere: unpinned).  Same tolerances as tests/test_gpu_parity.py / tests/test_gpu_ragged.py."""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu

CASES = [("iso", 10, {}), ("thermal", 10, dict(temperature=True)), ("sei", 10, dict(aging=True))]
IDS = [c[0] for c in CASES]


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


def _make(P, nr, opt, cathode="LCO"):
    p = P.petlion(cathode, Fickian_method="spectral", **{k: ("SEI" if k == "aging" else v) for k, v in opt.items()})
    m = O.make_model(cathode, Fickian_method="spectral", **opt)
    return p, m


@pytest.mark.parametrize("family,nr,opt", CASES, ids=IDS)
def test_pattern_and_resjac(P, family, nr, opt):
    p, m = _make(P, nr, opt)
    L = O.layout(m)
    assert p.N.tot == L.N_tot == 2 * 30 + (nr + 2) * 20 + 1 + (50 if family == "thermal" else 0) + (21 if family == "sei" else 0)
    for method in ("I", "V", "P"):
        cp, rv = O.jac_pattern(m, method)
        cp2, rv2 = p.jac_pattern(method)
        assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
    # 27 more entries per particle than the finite-difference scheme: the dense block (100 vs 82) and the j column (10 vs 1)
    assert cp[-1] - O.jac_pattern(O.make_model("LCO", **opt), "P")[0][-1] == 27 * 20
    B = 6
    tho = util.oracle_theta_batch(B, first=900)
    th = util.product_theta_from_oracle(p, tho)
    Y, YP = util.random_states(m, tho, seed=5)
    gam = np.array([0.01, 0.2, 1.0, 5.0, 50.0, 0.05])
    res, nz = p.resjac(Y, YP, gam, method="I", value=1.0, theta=th)
    run = O.make_run("I", 1.0)
    cp, rv = O.jac_pattern(m, "I")
    for s in range(B):
        r_ref = O.residual(m, tho[s], run, 0.0, Y[s], YP[s])
        j_ref = O.jacobian(m, tho[s], run, 0.0, Y[s], YP[s], gam[s])
        scale = np.zeros(L.N_tot)
        for c in range(L.N_tot):
            k = slice(cp[c], cp[c + 1])
            np.maximum.at(scale, rv[k], np.abs(j_ref[k]) * max(abs(Y[s][c]), 1e-12))
        scale = np.maximum(scale, np.abs(r_ref))
        tol = 1e-9 if family == "thermal" else 1e-10
        mask = np.ones(L.N_tot, bool)
        if family == "thermal":          # T rows: the reference's A_tot*T form carries ~1e-5 K/s of cancellation noise
            mask[L.T:L.T + 50] = False
        assert np.all(np.abs(res[s] - r_ref)[mask] <= tol * scale[mask] + 1e-300)
        rowmax = np.zeros(L.N_tot)
        np.maximum.at(rowmax, rv, np.abs(j_ref))
        assert np.all(np.abs(nz[s] - j_ref) <= 1e-6 * rowmax[rv] + 1e-300)


@pytest.mark.parametrize("family,nr,opt", CASES, ids=IDS)
def test_linear_solve_equals_dense(P, family, nr, opt):
    p, m = _make(P, nr, opt)
    B = 4
    tho = util.oracle_theta_batch(B, first=40)
    th = util.product_theta_from_oracle(p, tho)
    Y, YP = util.random_states(m, tho, seed=3)
    gam = np.array([50.0, 1.0, 0.05, 0.01])
    for method, value in (("I", 2.0), ("V", 4.0)):
        run = O.make_run(method, value)
        cp, rv = O.jac_pattern(m, method)
        N = len(cp) - 1
        rng = np.random.default_rng(2)
        Js, rhs = [], []
        for s in range(B):
            nzv = O.jacobian(m, tho[s], run, 0.0, Y[s], YP[s], gam[s])
            J = np.zeros((N, N))
            for c in range(N):
                J[rv[cp[c]:cp[c + 1]], c] = nzv[cp[c]:cp[c + 1]]
            Js.append(J); rhs.append(rng.normal(size=N) * np.abs(J).max(axis=1) * 1e-3)
        rhs = np.stack(rhs)
        x, st = p.linear_solve(Y, YP, gam, rhs, method=method, value=value, theta=th)
        for s in range(B):
            xr = np.linalg.solve(Js[s], rhs[s])
            rr = np.linalg.norm(Js[s] @ x[s] - rhs[s]) / np.linalg.norm(rhs[s])
            rr_ref = np.linalg.norm(Js[s] @ xr - rhs[s]) / np.linalg.norm(rhs[s])
            assert rr < 50 * rr_ref + 1e-11, (method, s, rr, rr_ref)


@pytest.mark.parametrize("family,nr,opt", CASES, ids=IDS)
def test_newton_and_cccv(P, family, nr, opt):
    p, m = _make(P, nr, opt)
    L = O.layout(m)
    B = 8
    tho = util.oracle_theta_batch(B, first=300)
    th = util.product_theta_from_oracle(p, tho)
    util.set_theta_batch(p, th)
    soc = np.linspace(0.1, 0.9, B)
    cur = np.where(np.arange(B) % 2 == 0, -1.0, 2.0)
    Y0 = p.initial_guess(soc, theta=th)
    Y0[:, L.I] = cur
    st, Y, YP = p.newton_init(Y0, method="I", value=cur, theta=th)
    opts = O.default_opts()
    for s in range(B):
        it, y, yp = O.newton_init(m, tho[s], O.make_run("I", cur[s]), opts, Y0[s])
        assert st[s] == it
        np.testing.assert_allclose(Y[s], y, rtol=1e-9, atol=1e-12)
    # integrator level: a 1.5C charge to a voltage bound, then the CV hold (simulate!)
    sol = P.simulate(p, I=1.5, SOC=0.1, V_max=4.05)
    b = O.default_bounds("LCO", V_max=4.05)
    ref = O.simulate_batch(m, tho, O.make_run("I", 1.5), opts, b, SOC0=0.1, n_save_max=512, nthreads=8)
    s = sol.results[-1].summary
    same = s["n_steps"] == ref["n_steps"]
    assert np.mean(same) >= 0.6, (s["n_steps"], ref["n_steps"])
    assert np.array_equal(s["flag"], ref["flag"]) and np.all(ref["flag"] == 2)
    # (thermal: equal step counts are not equal step times -- the conduction rows' cancellation noise enters the error norm
    #  and with it the step sizes; same bound as tests/test_gpu_thermal.py::_compare)
    np.testing.assert_allclose(s["t_end"][same], ref["t_end"][same], rtol=1e-4 if family == "thermal" else 1e-6)
    np.testing.assert_allclose(s["V_end"][same], ref["V_end"][same], rtol=1e-6)
    np.testing.assert_allclose(s["t_end"], ref["t_end"], rtol=2e-3)
    np.testing.assert_allclose(s["SOC_end"], ref["SOC_end"], rtol=2e-3)
    P.simulate_(sol, p, 600, V="hold", V_max=4.05)
    ref2 = O.simulate_batch(m, tho, O.make_run("V", 0.0, tf=600, input_kind="hold", new_run=False), opts, b,
                            state=ref["state"], n_save_max=512, nthreads=8)
    s2 = sol.results[-1].summary
    assert np.array_equal(s2["flag"], ref2["flag"])
    np.testing.assert_allclose(s2["I_end"], ref2["I_end"], rtol=5e-3, atol=1e-6)
    np.testing.assert_allclose(s2["SOC_end"], ref2["SOC_end"], rtol=2e-3)


def test_spectral_discharge_tight_and_nmc(P):
    """1C discharges to the exit at reltol = abstol = 1e-9 (V at fixed times through both dense outputs, rtol 1e-6); the NMC
    parameter set; and the scheme against the finite-difference one: the same cell, a few mV apart at the end of a discharge."""
    p, m = _make(P, 10, {})
    B = 16
    tho = util.oracle_theta_batch(B, first=77)
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    tol = 1e-9
    dense_t = np.linspace(5.0, 3400.0, 80)
    sol = P.simulate(p, I=-1, SOC=1, n_save_max=0, dense_t=dense_t, reltol=tol, abstol=tol, reltol_init=tol, abstol_init=tol)
    o = O.default_opts(reltol=tol, abstol=tol, reltol_init=tol, abstol_init=tol)
    ref = O.simulate_batch(m, tho, O.make_run("I", -1.0), o, O.default_bounds("LCO"), SOC0=1.0, nthreads=8, dense_t=dense_t)
    s = sol.results[-1].summary
    assert np.array_equal(s["flag"], ref["flag"])
    np.testing.assert_allclose(s["t_end"], ref["t_end"], rtol=5e-4)      # the exit blend is h^2-accurate (tests/test_gpu_tight.py)
    both = ~np.isnan(sol.dense["V"]) & ~np.isnan(ref["dense"]["V"])
    assert both.mean() > 0.9
    np.testing.assert_allclose(sol.dense["V"][both], ref["dense"]["V"][both], rtol=1e-6)
    np.testing.assert_allclose(sol.dense["SOC"][both], ref["dense"]["SOC"][both], atol=1e-6)
    pf = P.petlion("LCO")
    util.set_theta_batch(pf, util.product_theta_from_oracle(pf, tho))
    solf = P.simulate(pf, I=-1, SOC=1, n_save_max=0, dense_t=dense_t, reltol=tol, abstol=tol, reltol_init=tol, abstol_init=tol)
    bothf = both & ~np.isnan(solf.dense["V"])
    dV = np.abs(sol.dense["V"][bothf] - solf.dense["V"][bothf])
    assert 1e-7 < dV.max() < 0.02, dV.max()
    pn, mn = _make(P, 10, {}, cathode="NMC")
    thn = util.oracle_theta_batch(4, cathode="NMC")
    util.set_theta_batch(pn, util.product_theta_from_oracle(pn, thn))
    soln = P.simulate(pn, 1800, I=1, SOC=0)
    refn = O.simulate_batch(mn, thn, O.make_run("I", 1.0, tf=1800), O.default_opts(), O.default_bounds("NMC"), SOC0=0.0, nthreads=4)
    sn = soln.results[-1].summary
    assert np.array_equal(sn["flag"], refn["flag"])
    same = sn["n_steps"] == refn["n_steps"]
    assert same.mean() >= 0.5
    np.testing.assert_allclose(sn["V_end"][same], refn["V_end"][same], rtol=1e-6)
    np.testing.assert_allclose(sn["V_end"], refn["V_end"], rtol=1e-3)
