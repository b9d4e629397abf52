"""GPU parity tests of the temperature = true variant (351 DAEs, BASELINE.json configs[2]) through the
C ABI, against the CPU oracle (itself pinned to the reference's thermal notebook, tests/test_oracle_golden.py).

Tolerances (FP64): residual 1e-9 of the row scale on physical states; Jacobian entries 1e-9 of the row
maximum and 1e-6 entry-wise; structured linear solve: LAPACK-level residual; trajectories: identical
step sequences, V / SOC / T within rtol 1e-6.
"""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu

T_BOUNDS = dict(T_max=40 + 273.15, V_max=4.1, I_max=4.0, I_min=1 / 20)


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


@pytest.fixture(scope="module")
def lcoT(P):
    return P.petlion("LCO", temperature=True)


@pytest.fixture(scope="module")
def mT():
    return O.make_model("LCO", temperature=True)


def _groups(L):
    return [("c_e", L.c_e, L.c_s_p), ("c_s", L.c_s_p, L.T), ("T", L.T, L.j), ("j", L.j, L.phi_e),
            ("phi_e", L.phi_e, L.phi_s), ("phi_s", L.phi_s, L.I), ("I", L.I, L.I + 1)]


def _group_of(L, i):
    for name, a, b in _groups(L):
        if a <= i < b:
            return f"{name}[{i - a}]"
    return str(i)


def physical_states(m, tho, t_mid=120.0, I=4.0):
    """(Y, Y') of an oracle run stopped at t_mid: smooth temperature field, consistent algebraic states"""
    b = O.default_bounds("LCO", **T_BOUNDS)
    r = O.simulate_batch(m, tho, O.make_run("I", I, tf=t_mid), O.default_opts(), b, SOC0=0.0, nthreads=8)
    assert np.all(r["flag"] == 0)
    return r["state"]["Y"], r["state"]["YP"]


def test_theta_keys_thermal(lcoT):
    assert lcoT.θ_keys == sorted(lcoT.θ_keys)
    assert len(lcoT.θ_keys) == 56
    od = O.theta_dict("LCO")
    th = util.product_theta_from_oracle(lcoT, np.array([list(od.values())]))
    assert np.array_equal(th[0], np.array(list(lcoT.θ.values())))
    assert (lcoT.N.tot, lcoT.N.diff) == (351, 280)


@pytest.mark.parametrize("method", ["I", "V", "P"])
def test_jac_pattern_equals_oracle(lcoT, mT, method):
    cp, rv = O.jac_pattern(mT, method)
    cp2, rv2 = lcoT.jac_pattern(method)
    L = O.layout(mT)
    ref = {(int(r), c) for c in range(L.N_tot) for r in rv[cp[c]:cp[c + 1]]}
    got = {(int(r), c) for c in range(L.N_tot) for r in rv2[cp2[c]:cp2[c + 1]]}
    missing = sorted(ref - got)[:10]
    extra = sorted(got - ref)[:10]
    assert not missing and not extra, ("missing", [(_group_of(L, r), _group_of(L, c)) for r, c in missing],
                                       "extra", [(_group_of(L, r), _group_of(L, c)) for r, c in extra])
    assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
    assert len(rv2) == (2883 if method == "I" else 2884 if method == "V" else 2885)


def test_initial_guess(lcoT, mT):
    tho = util.oracle_theta_batch(8)
    tho[4:, O.theta_names().index("T0")] = 301.0
    th = util.product_theta_from_oracle(lcoT, tho)
    soc = np.linspace(0, 1, 8)
    Y0 = lcoT.initial_guess(soc, theta=th)
    for s in range(8):
        np.testing.assert_allclose(Y0[s], O.initial_guess(mT, tho[s], soc[s]), rtol=1e-13, atol=0)


def _check_resjac(lcoT, mT, tho, Y, YP, gam, method, value, res_tol, what):
    L = O.layout(mT)
    N = L.N_tot
    th = util.product_theta_from_oracle(lcoT, tho)
    res, nz = lcoT.resjac(Y, YP, gam, method=method, value=value, theta=th)
    run = O.make_run(method, value)
    cp, rv = O.jac_pattern(mT, method)
    cols = np.repeat(np.arange(N), np.diff(cp))
    worst_r, worst_j, worst_e = (0.0, None), (0.0, None), (0.0, None)
    for s in range(len(tho)):
        r_ref = O.residual(mT, tho[s], run, 0.0, Y[s], YP[s])
        j_ref = O.jacobian(mT, tho[s], run, 0.0, Y[s], YP[s], gam[s])
        scale = np.zeros(N)
        np.maximum.at(scale, rv, np.abs(j_ref) * np.maximum(np.abs(Y[s][cols]), 1e-12))
        scale = np.maximum(scale, np.abs(r_ref))
        er = np.abs(res[s] - r_ref) / (scale + 1e-300)
        if er.max() > worst_r[0]:
            worst_r = (float(er.max()), (s, _group_of(L, int(er.argmax())), float(res[s][er.argmax()]), float(r_ref[er.argmax()])))
        rowmax = np.zeros(N)
        np.maximum.at(rowmax, rv, np.abs(j_ref))
        ej = np.abs(nz[s] - j_ref) / rowmax[rv]
        if ej.max() > worst_j[0]:
            k = int(ej.argmax())
            worst_j = (float(ej.max()), (s, _group_of(L, int(rv[k])), _group_of(L, int(cols[k])), float(nz[s][k]), float(j_ref[k])))
        ee = np.abs(nz[s] - j_ref) / np.maximum(np.abs(j_ref), 1e-9 * rowmax[rv])
        if ee.max() > worst_e[0]:
            k = int(ee.argmax())
            worst_e = (float(ee.max()), (s, _group_of(L, int(rv[k])), _group_of(L, int(cols[k])), float(nz[s][k]), float(j_ref[k])))
    print(what, method, "residual", worst_r, "jac(rowmax)", worst_j, "jac(entry)", worst_e)
    assert worst_r[0] < res_tol, worst_r
    assert worst_j[0] < 1e-9, worst_j
    assert worst_e[0] < 1e-6, worst_e


@pytest.mark.parametrize("method,value", [("I", 4.0), ("V", 4.0), ("P", 300.0)])
def test_resjac_parity_physical_states(lcoT, mT, method, value):
    B = 12
    tho = util.oracle_theta_batch(B)
    Y, YP = physical_states(mT, tho)
    gam = np.random.default_rng(5).uniform(0.01, 50.0, size=B)
    _check_resjac(lcoT, mT, tho, Y, YP, gam, method, value, 1e-9, "physical")


def test_resjac_parity_random_states(lcoT, mT):
    B = 12
    tho = util.oracle_theta_batch(B, first=100)
    Y, YP = util.random_states(mT, tho, seed=3)
    gam = np.random.default_rng(6).uniform(0.01, 50.0, size=B)
    _check_resjac(lcoT, mT, tho, Y, YP, gam, "I", 1.0, 1e-9, "random")


def _dense(m, tho_row, run, Y, YP, gam):
    cp, rv = O.jac_pattern(m, ["I", "V", "P"][run.method])
    nz = O.jacobian(m, tho_row, run, 0.0, Y, YP, gam)
    N = len(cp) - 1
    J = np.zeros((N, N))
    for c in range(N):
        J[rv[cp[c]:cp[c + 1]], c] = nz[cp[c]:cp[c + 1]]
    return J


@pytest.mark.parametrize("method,value", [("I", 4.0), ("V", 4.0), ("P", 300.0)])
def test_linear_solve_equals_dense(lcoT, mT, method, value):
    """the structured factorisation (eigen-basis particles, collector chains, twisted 4x4 block-Thomas with
    the two-node-wide T couplings, current border) against LAPACK on the oracle's Jacobian"""
    B = 9
    tho = util.oracle_theta_batch(B, first=40)
    th = util.product_theta_from_oracle(lcoT, tho)
    Y, YP = physical_states(mT, tho)
    gam = np.array([50.0, 5.0, 0.5, 0.05, 0.01, 20.0, 2.0, 0.2, 0.02])
    rng = np.random.default_rng(2)
    run = O.make_run(method, value)
    Js = [_dense(mT, tho[s], run, Y[s], YP[s], gam[s]) for s in range(B)]
    rhs = np.stack([rng.normal(size=J.shape[0]) * np.abs(J).max(axis=1) * 1e-3 for J in Js])
    x, st = lcoT.linear_solve(Y, YP, gam, rhs, method=method, value=value, theta=th)
    assert np.all(st == 0)
    L = O.layout(mT)
    for s in range(B):
        xr = np.linalg.solve(Js[s], rhs[s])
        rr = np.linalg.norm(Js[s] @ x[s] - rhs[s]) / np.linalg.norm(rhs[s])
        rr_ref = np.linalg.norm(Js[s] @ xr - rhs[s]) / np.linalg.norm(rhs[s])
        err = np.abs(x[s] - xr) / np.max(np.abs(xr))
        print(method, s, "relres", rr, "lapack", rr_ref, "max err", err.max(), _group_of(L, int(err.argmax())))
        assert rr < 10 * rr_ref + 1e-12, (s, rr, rr_ref, _group_of(L, int(err.argmax())))
        assert err.max() < 1e-5, (s, err.max(), _group_of(L, int(err.argmax())))


def test_linear_solve_isothermal(P):
    p = P.petlion("LCO")
    m = O.make_model("LCO")
    B = 6
    tho = util.oracle_theta_batch(B, first=40)
    th = util.product_theta_from_oracle(p, tho)
    r = O.simulate_batch(m, tho, O.make_run("I", -1.0, tf=900.0), O.default_opts(), O.default_bounds("LCO"), SOC0=1.0)
    Y, YP = r["state"]["Y"], r["state"]["YP"]
    gam = np.array([50.0, 5.0, 0.5, 0.05, 0.01, 1.0])
    run = O.make_run("I", -1.0)
    Js = [_dense(m, tho[s], run, Y[s], YP[s], gam[s]) for s in range(B)]
    rng = np.random.default_rng(2)
    rhs = np.stack([rng.normal(size=301) * np.abs(J).max(axis=1) * 1e-3 for J in Js])
    x, st = p.linear_solve(Y, YP, gam, rhs, method="I", value=-1.0, theta=th)
    for s in range(B):
        xr = np.linalg.solve(Js[s], rhs[s])
        rr = np.linalg.norm(Js[s] @ x[s] - rhs[s]) / np.linalg.norm(rhs[s])
        rr_ref = np.linalg.norm(Js[s] @ xr - rhs[s]) / np.linalg.norm(rhs[s])
        assert rr < 10 * rr_ref + 1e-12, (s, rr, rr_ref)


def test_newton_init_parity(lcoT, mT):
    L = O.layout(mT)
    B = 16
    tho = util.oracle_theta_batch(B)
    th = util.product_theta_from_oracle(lcoT, tho)
    soc = np.linspace(0.05, 0.95, B)
    cur = np.where(np.arange(B) % 2 == 0, -1.0, 4.0)
    Y0 = lcoT.initial_guess(soc, theta=th)
    Y0[:, L.I] = cur
    st, Y, YP = lcoT.newton_init(Y0, method="I", value=cur, theta=th)
    opts = O.default_opts()
    for s in range(B):
        it, y, yp = O.newton_init(mT, tho[s], O.make_run("I", cur[s]), opts, Y0[s])
        assert st[s] == it
        np.testing.assert_allclose(Y[s], y, rtol=1e-9, atol=1e-12)
        scale = np.maximum(np.abs(yp), 1e-6 * np.abs(yp).max())
        # T rows: the reference (and the oracle) evaluate conduction as A_tot*T, a sum of terms of size
        # lambda/h^2*T/(rho Cp) ~ 1e10 K/s that cancel to ~0: its own round-off floor is ~1e-5 K/s
        # (the CUDA path differences T first and returns the exact 4e-12 K/s Joule term at uniform T)
        scale[L.T:L.j] = np.maximum(scale[L.T:L.j], 20.0)
        bad = int(np.argmax(np.abs(YP[s] - yp) / scale))
        assert np.max(np.abs(YP[s] - yp) / scale) < 1e-6, (s, _group_of(L, bad), YP[s][bad], yp[bad])


def _compare(sol, ref, rtol=1e-6, min_identical=1.0):
    summ = sol.results[-1].summary
    same = summ["n_steps"] == ref["n_steps"]
    print("identical step counts:", float(np.mean(same)), "gpu", summ["n_steps"][:8], "ref", ref["n_steps"][:8],
          "flags", summ["flag"][:8], ref["flag"][:8])
    assert np.mean(same) >= min_identical, (np.mean(same), summ["n_steps"], ref["n_steps"], summ["flag"], ref["flag"])
    idx = np.where(same)[0]
    assert np.array_equal(summ["flag"][idx], ref["flag"][idx])
    np.testing.assert_allclose(summ["t_end"][idx], ref["t_end"][idx], rtol=100 * rtol)
    np.testing.assert_allclose(summ["V_end"][idx], ref["V_end"][idx], rtol=rtol)
    np.testing.assert_allclose(summ["SOC_end"][idx], ref["SOC_end"][idx], rtol=100 * rtol, atol=1e-8)
    np.testing.assert_allclose(summ["T_end"][idx], ref["T_end"][idx], rtol=rtol)
    n_exact = 0
    for s in idx:
        n = ref["traj_n"][s]
        tg, tr = sol.t[s, :n], ref["traj"]["t"][s, :n]
        # Same step/order decisions.  After a step-size change by the continuous factor
        # rr = (2 err + 1e-4)^(-1/(k+1)) the time grid inherits the relative round-off of the error estimate
        # (a difference of nearly equal vectors; the reference's own T rows carry ~1e-5 K/s of cancellation
        # noise), and the last point is the back-interpolation onto the bound (checks.jl:37-41).  Systems whose
        # grid is identical are compared at the north-star rtol 1e-6 on V, SOC and T; on the few whose grid
        # drifted (observed <= 1.5e-5 relative) V is compared with the slope * time-shift allowance.
        exact = np.allclose(tg[:n - 1], tr[:n - 1], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(tg, tr, rtol=100 * rtol, atol=1e-9)
        Vg, Vr = sol.V[s, :n], ref["traj"]["V"][s, :n]
        if exact:
            n_exact += 1
            np.testing.assert_allclose(Vg[:n - 1], Vr[:n - 1], rtol=rtol)
            np.testing.assert_allclose(sol.SOC[s, :n - 1], ref["traj"]["SOC"][s, :n - 1], rtol=rtol, atol=1e-8)
        slope = np.abs(np.gradient(Vr, np.maximum(tr, 1e-12) + np.arange(n) * 1e-12))
        assert np.all(np.abs(Vg - Vr) <= rtol * np.abs(Vr) + 2 * slope * np.abs(tg - tr) + 1e-12)
        np.testing.assert_allclose(sol.SOC[s, :n], ref["traj"]["SOC"][s, :n], rtol=100 * rtol, atol=1e-8)
    print("grid-identical systems:", n_exact, "of", len(idx))
    assert n_exact >= len(idx) // 4      # (6-8 of 12 with most builds, 5 with one: which systems stay on one grid through three
                                         # segments is round-off luck; all of them are compared above at their own tolerance)


def test_simulate_thermal_4C_matches_reference_notebook(P, lcoT, mT, goldens):
    """examples/fast_charging_CC-CT-CV.ipynb cell 7: simulate(p, I=4) from SOC 0 until T_max = 40 C"""
    for k, v in zip(lcoT.θ_keys, P.petlion("LCO", temperature=True).θ.values()):
        lcoT.θ[k] = v
    sol = P.simulate(lcoT, I=4, SOC=0, **T_BOUNDS)
    s = sol.results[-1].summary
    g = goldens["summaries"]["thermal_4C_to_Tmax"]
    assert s["flag"][0] == 5 and sol.results[-1].exit_reason[0] == "Above max. temperature"
    assert round(float(s["t_end"][0]), 2) == g["t_s"]
    assert round(float(s["V_end"][0]), 4) == g["V"]
    assert round(float(s["SOC_end"][0]), 4) == g["SOC"]
    assert round(float(s["T_end"][0]) - 273.15, 4) == g["T_C"]
    assert round(float(s["V_end"][0] * 4.0 * lcoT.I1C(1)[0]), 2) == g["P"]
    lad = np.array(goldens["ladder_4C_thermal"]["t"])
    assert s["n_steps"][0] == len(lad) - 1 == 76
    assert np.all(np.abs(sol.t[0, :77] - lad) <= 0.002 + 1e-3 * lad)
    ref = O.simulate_batch(mT, O.theta_defaults("LCO"), O.make_run("I", 4.0), O.default_opts(),
                           O.default_bounds("LCO", **T_BOUNDS), SOC0=0.0, n_save_max=512)
    _compare(sol, ref)


def test_simulate_thermal_batch_cc_cv(P, lcoT, mT):
    """configs[2] at test size: randomised batch, 4C CC charge to 4.1 V (T_max at its default 55 C), then
    V = :hold until SOC_max / I_min, state handed over on the device side of the ABI"""
    B = 48
    tho = util.oracle_theta_batch(B, first=2000)
    th = util.product_theta_from_oracle(lcoT, tho)
    util.set_theta_batch(lcoT, th)
    sol = P.simulate(lcoT, I=4, SOC=0, V_max=4.1)
    b = O.default_bounds("LCO", V_max=4.1)
    ref = O.simulate_batch(mT, tho, O.make_run("I", 4.0), O.default_opts(), b, SOC0=0.0, n_save_max=512, nthreads=8)
    _compare(sol, ref, min_identical=0.9)
    # temperature trajectory of the first system
    n = ref["traj_n"][0]
    assert np.all(np.isfinite(sol.T[0, :n])) and sol.T[0, n - 1] > sol.T[0, 0]
    P.simulate_(sol, lcoT, V="hold", V_max=4.1)
    ref2 = O.simulate_batch(mT, tho, O.make_run("V", 0.0, input_kind="hold", new_run=False), O.default_opts(), b,
                            state=ref["state"], n_save_max=512, nthreads=8)
    s2 = sol.results[-1].summary
    same = s2["n_steps"] == ref2["n_steps"]
    print("CV identical", float(np.mean(same)), s2["flag"][:8], ref2["flag"][:8])
    assert np.mean(same) >= 0.8
    assert np.array_equal(s2["flag"][same], ref2["flag"][same])
    np.testing.assert_allclose(s2["t_end"][same], ref2["t_end"][same], rtol=1e-6)
    # (systems that run to tf = 1e6 s end with I ~ 1e-10 C: absolute tolerance)
    np.testing.assert_allclose(s2["I_end"][same], ref2["I_end"][same], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(s2["T_end"][same], ref2["T_end"][same], rtol=1e-6)


def test_thermal_discharge_and_tight_tolerance(P, lcoT, mT):
    B = 6
    tho = util.oracle_theta_batch(B, first=7000)
    th = util.product_theta_from_oracle(lcoT, tho)
    util.set_theta_batch(lcoT, th)
    sol = P.simulate(lcoT, 900.0, I=-2, SOC=1, abstol=1e-9, reltol=1e-9, n_save_max=8)
    ref = O.simulate_batch(mT, tho, O.make_run("I", -2.0, tf=900.0), O.default_opts(abstol=1e-9, reltol=1e-9,
                           abstol_init=1e-9, reltol_init=1e-9), O.default_bounds("LCO"), SOC0=1.0, nthreads=8)
    s = sol.results[-1].summary
    np.testing.assert_allclose(s["V_end"], ref["V_end"], rtol=1e-6)
    np.testing.assert_allclose(s["T_end"], ref["T_end"], rtol=1e-7)


# ---------------------------------------------------------------------------------------------------
# dT = :hold (constant-temperature control, input_methods.jl:182-189) and the README's CC-CT-CV fast charge
# ---------------------------------------------------------------------------------------------------
def test_dT_pattern_and_resjac(lcoT, mT):
    cp, rv = O.jac_pattern(mT, "dT")
    cp2, rv2 = lcoT.jac_pattern("dT")
    assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
    B = 8
    tho = util.oracle_theta_batch(B, first=300)
    Y, YP = physical_states(mT, tho)
    gam = np.random.default_rng(9).uniform(0.01, 50.0, size=B)
    _check_resjac(lcoT, mT, tho, Y, YP, gam, "dT", 0.0, 1e-9, "physical")


def test_dT_linear_solve_equals_dense(lcoT, mT):
    B = 5
    tho = util.oracle_theta_batch(B, first=60)
    th = util.product_theta_from_oracle(lcoT, tho)
    Y, YP = physical_states(mT, tho)
    gam = np.array([50.0, 5.0, 0.5, 0.05, 2.0])
    run = O.make_run("dT", 0.0)
    cp, rv = O.jac_pattern(mT, "dT")
    N = len(cp) - 1
    rng = np.random.default_rng(4)
    Js, rhs = [], []
    for s in range(B):
        nz = O.jacobian(mT, tho[s], run, 0.0, Y[s], YP[s], gam[s])
        J = np.zeros((N, N))
        for c in range(N):
            J[rv[cp[c]:cp[c + 1]], c] = nz[cp[c]:cp[c + 1]]
        Js.append(J); rhs.append(rng.normal(size=N) * np.abs(J).max(axis=1) * 1e-3)
    rhs = np.stack(rhs)
    x, st = lcoT.linear_solve(Y, YP, gam, rhs, method="dT", value=0.0, theta=th)
    for s in range(B):
        xr = np.linalg.solve(Js[s], rhs[s])
        rr = np.linalg.norm(Js[s] @ x[s] - rhs[s]) / np.linalg.norm(rhs[s])
        rr_ref = np.linalg.norm(Js[s] @ xr - rhs[s]) / np.linalg.norm(rhs[s])
        print("dT solve", s, rr, rr_ref)
        # the control row has a zero corner and the temperature response to the current is tiny, so the border
        # (Schur complement) step amplifies round-off more than LAPACK's pivoted LU; far below Newton's needs
        # (grows like 1/gamma: 2e-9 at gamma = 0.05)
        assert rr < max(100 * rr_ref, 1e-7), (s, rr, rr_ref)


def test_dT_newton_init_parity(lcoT, mT):
    """the algebraic re-initialisation when the control switches from I = 4C to dT = 0: I drops to ~3.3C"""
    L = O.layout(mT)
    B = 8
    tho = util.oracle_theta_batch(B, first=500)
    th = util.product_theta_from_oracle(lcoT, tho)
    Y0, _ = physical_states(mT, tho, t_mid=200.0)
    st, Y, YP = lcoT.newton_init(Y0, method="dT", value=0.0, theta=th)
    for s in range(B):
        it, y, yp = O.newton_init(mT, tho[s], O.make_run("dT", 0.0), O.default_opts(), Y0[s])
        assert st[s] == it, (s, st[s], it)
        # the root of  sum_x w_x rhs_T[x](I) = 0  is only defined down to the noise of rhs_T: the reference (and
        # the oracle) evaluate conduction as A_tot*T, ~1e10-sized terms cancelling to ~1e-2 K/s, i.e. ~1e-7 of
        # noise in the control row and ~1e-6 in I (the CUDA path differences temperatures first)
        np.testing.assert_allclose(Y[s], y, rtol=5e-6, atol=1e-12)
        assert abs(Y[s][L.I] - 4.0) > 0.1


def test_cc_ct_cv_fast_charge(P, lcoT, mT, goldens):
    """README.md:23-37 / examples/fast_charging_CC-CT-CV.ipynb: simulate(I=4) -> simulate!(dT=:hold) ->
    simulate!(V=:hold), nominal cell plus a small randomised batch, against the oracle; the nominal cell also
    against the notebook's printed summaries (3 digits: see tests/test_oracle_golden.py for why not more)."""
    B = 12
    tho = util.oracle_theta_batch(B, first=4000)
    tho[0] = O.theta_defaults("LCO")
    th = util.product_theta_from_oracle(lcoT, tho)
    util.set_theta_batch(lcoT, th)
    b = O.default_bounds("LCO", **T_BOUNDS)
    sol = P.simulate(lcoT, I=4, SOC=0, **T_BOUNDS)
    r1 = O.simulate_batch(mT, tho, O.make_run("I", 4.0), O.default_opts(), b, SOC0=0.0, n_save_max=512, nthreads=8)
    _compare(sol, r1, min_identical=0.9)
    P.simulate_(sol, lcoT, dT="hold", **T_BOUNDS)
    r2 = O.simulate_batch(mT, tho, O.make_run("dT", 0.0, input_kind="hold", new_run=False), O.default_opts(), b,
                          state=r1["state"], n_save_max=512, nthreads=8)
    s2 = sol.results[-1].summary
    print("CT: gpu steps", s2["n_steps"], "ref", r2["n_steps"], "flags", s2["flag"], r2["flag"])
    same = (s2["n_steps"] == r2["n_steps"]) & (s2["flag"] == r2["flag"])
    assert np.mean(same) >= 0.6
    np.testing.assert_allclose(s2["t_end"][same], r2["t_end"][same], rtol=1e-5)
    np.testing.assert_allclose(s2["I_end"][same], r2["I_end"][same], rtol=1e-5)
    np.testing.assert_allclose(s2["T_end"][same], r2["T_end"][same], rtol=1e-6)
    # everybody, including the systems whose step sequence flipped somewhere: integration-tolerance agreement
    np.testing.assert_allclose(s2["t_end"], r2["t_end"], rtol=5e-3)
    np.testing.assert_allclose(s2["I_end"], r2["I_end"], rtol=5e-3)
    g = goldens["summaries"]["thermal_CT_hold"]
    assert s2["t_end"][0] == pytest.approx(g["t_s"], rel=1e-3) and s2["I_end"][0] == pytest.approx(g["I_C"], rel=1e-3)
    assert s2["T_end"][0] - 273.15 == pytest.approx(40.0, abs=1e-3)
    P.simulate_(sol, lcoT, V="hold", **T_BOUNDS)
    s3 = sol.results[-1].summary
    g3 = goldens["summaries"]["thermal_CV_after_CT"]
    assert s3["flag"][0] == 4 and sol.results[-1].exit_reason[0] == "Above max. SOC"
    assert s3["t_end"][0] == pytest.approx(g3["t_s"], rel=1e-2)
    assert s3["I_end"][0] == pytest.approx(g3["I_C"], rel=3e-2)
    assert s3["T_end"][0] - 273.15 == pytest.approx(g3["T_C"], abs=0.05)
    assert len(sol.results) == 3
