"""GPU parity tests of the aging = :SEI family (film, SOH, j_s; N = 322 for N = (10,10,10)) through the C ABI
against the CPU oracle.  The reference ships no executed SEI example, so the SEI rows of the oracle are a
restatement of residuals.jl:260-297, 519-552 without a golden pin (the rest of the model is pinned)."""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


@pytest.fixture(scope="module")
def lcoS(P):
    return P.petlion("LCO", aging="SEI")


@pytest.fixture(scope="module")
def mS():
    return O.make_model("LCO", aging=True)


def test_theta_keys_sei(lcoS):
    assert lcoS.θ_keys == sorted(lcoS.θ_keys)
    assert len(lcoS.θ_keys) == 42
    for k in ("M_n", "R_SEI", "Uref_s", "i_0_jside", "k_n_aging", "w", "ρ_n"):
        assert k in lcoS.θ_keys
    od = O.theta_dict("LCO")
    th = util.product_theta_from_oracle(lcoS, np.array([list(od.values())]))
    assert np.array_equal(th[0], np.array(list(lcoS.θ.values())))
    assert (lcoS.N.tot, lcoS.N.diff) == (322, 241)


@pytest.mark.parametrize("method", ["I", "V", "P"])
def test_jac_pattern_equals_oracle(lcoS, mS, method):
    cp, rv = O.jac_pattern(mS, method)
    cp2, rv2 = lcoS.jac_pattern(method)
    assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
    assert len(rv2) == {"I": 2269, "V": 2270, "P": 2271}[method]


def _states(mS, tho, cur, soc0, t_mid):
    b = O.default_bounds("LCO", V_max=4.3)
    r = O.simulate_batch(mS, tho, O.make_run("I", cur, tf=t_mid), O.default_opts(), b, SOC0=soc0, nthreads=8)
    assert np.all(r["flag"] == 0)
    return r["state"]["Y"], r["state"]["YP"]


@pytest.mark.parametrize("cur,soc0,method,value", [(1.0, 0.2, "I", 1.0), (2.0, 0.1, "V", 4.0), (-1.0, 0.9, "I", -1.0),
                                                   (1.5, 0.3, "P", 150.0)])
def test_resjac_parity(lcoS, mS, cur, soc0, method, value):
    B = 10
    L = O.layout(mS); N = L.N_tot
    tho = util.oracle_theta_batch(B, first=10)
    th = util.product_theta_from_oracle(lcoS, tho)
    Y, YP = _states(mS, tho, cur, soc0, 600.0)
    gam = np.random.default_rng(5).uniform(0.01, 50.0, size=B)
    res, nz = lcoS.resjac(Y, YP, gam, method=method, value=value, theta=th)
    run = O.make_run(method, value)
    cp, rv = O.jac_pattern(mS, method)
    cols = np.repeat(np.arange(N), np.diff(cp))
    for s in range(B):
        r_ref = O.residual(mS, tho[s], run, 0.0, Y[s], YP[s])
        j_ref = O.jacobian(mS, tho[s], run, 0.0, Y[s], YP[s], gam[s])
        scale = np.zeros(N)
        np.maximum.at(scale, rv, np.abs(j_ref) * np.maximum(np.abs(Y[s][cols]), 1e-12))
        scale = np.maximum(scale, np.abs(r_ref))
        er = np.abs(res[s] - r_ref) / (scale + 1e-300)
        assert er.max() < 1e-9, (s, int(er.argmax()), res[s][er.argmax()], r_ref[er.argmax()])
        rowmax = np.zeros(N); np.maximum.at(rowmax, rv, np.abs(j_ref))
        ej = np.abs(nz[s] - j_ref) / rowmax[rv]
        k = int(ej.argmax())
        assert ej.max() < 1e-9, (s, int(rv[k]), int(cols[k]), nz[s][k], j_ref[k])
        ee = np.abs(nz[s] - j_ref) / np.maximum(np.abs(j_ref), 1e-9 * rowmax[rv])
        k = int(ee.argmax())
        assert ee.max() < 1e-6, (s, int(rv[k]), int(cols[k]), nz[s][k], j_ref[k])
        if cur > 0:
            assert np.any(Y[s][L.j_s:L.j_s + 10] < 0)          # the side reaction is running


@pytest.mark.parametrize("cur,method,value", [(1.0, "I", 1.0), (2.0, "V", 4.0), (-1.0, "I", -1.0)])
def test_linear_solve_equals_dense(lcoS, mS, cur, method, value):
    B = 6
    tho = util.oracle_theta_batch(B, first=40)
    th = util.product_theta_from_oracle(lcoS, tho)
    Y, YP = _states(mS, tho, cur, 0.2 if cur > 0 else 0.9, 600.0)
    gam = np.array([50.0, 5.0, 0.5, 0.05, 0.01, 1.0])
    run = O.make_run(method, value)
    cp, rv = O.jac_pattern(mS, method)
    N = len(cp) - 1
    rng = np.random.default_rng(2)
    Js, rhs = [], []
    for s in range(B):
        nzv = O.jacobian(mS, tho[s], run, 0.0, Y[s], YP[s], gam[s])
        J = np.zeros((N, N))
        for c in range(N):
            J[rv[cp[c]:cp[c + 1]], c] = nzv[cp[c]:cp[c + 1]]
        Js.append(J); rhs.append(rng.normal(size=N) * np.abs(J).max(axis=1) * 1e-3)
    rhs = np.stack(rhs)
    x, st = lcoS.linear_solve(Y, YP, gam, rhs, method=method, value=value, theta=th)
    L = O.layout(mS)
    for s in range(B):
        xr = np.linalg.solve(Js[s], rhs[s])
        rr = np.linalg.norm(Js[s] @ x[s] - rhs[s]) / np.linalg.norm(rhs[s])
        rr_ref = np.linalg.norm(Js[s] @ xr - rhs[s]) / np.linalg.norm(rhs[s])
        err = np.abs(x[s] - xr) / np.max(np.abs(xr))
        print("sei solve", method, s, rr, rr_ref, err.max(), int(err.argmax()))
        assert rr < 20 * rr_ref + 1e-11, (s, rr, rr_ref, int(err.argmax()))


def test_newton_init_parity(lcoS, mS):
    L = O.layout(mS)
    B = 12
    tho = util.oracle_theta_batch(B)
    th = util.product_theta_from_oracle(lcoS, tho)
    soc = np.linspace(0.05, 0.95, B)
    cur = np.where(np.arange(B) % 2 == 0, -1.0, 2.0)
    Y0 = lcoS.initial_guess(soc, theta=th)
    for s in range(B):
        np.testing.assert_allclose(Y0[s], O.initial_guess(mS, tho[s], soc[s]), rtol=1e-13, atol=0)
    Y0[:, L.I] = cur
    st, Y, YP = lcoS.newton_init(Y0, method="I", value=cur, theta=th)
    opts = O.default_opts()
    for s in range(B):
        it, y, yp = O.newton_init(mS, tho[s], O.make_run("I", cur[s]), opts, Y0[s])
        assert st[s] == it
        np.testing.assert_allclose(Y[s], y, rtol=1e-9, atol=1e-14)
        scale = np.maximum(np.abs(yp), 1e-6 * np.abs(yp).max())
        assert np.max(np.abs(YP[s] - yp) / scale) < 1e-6


def _compare(sol, ref, rtol=1e-6, min_identical=0.9):
    summ = sol.results[-1].summary
    same = (summ["n_steps"] == ref["n_steps"]) & (summ["flag"] == ref["flag"])
    print("identical:", float(np.mean(same)), summ["n_steps"][:8], ref["n_steps"][:8], summ["flag"][:8], ref["flag"][:8])
    assert np.mean(same) >= min_identical
    idx = np.where(same)[0]
    np.testing.assert_allclose(summ["t_end"][idx], ref["t_end"][idx], rtol=5 * rtol)
    np.testing.assert_allclose(summ["V_end"][idx], ref["V_end"][idx], rtol=rtol)
    np.testing.assert_allclose(summ["SOC_end"][idx], ref["SOC_end"][idx], rtol=5 * rtol, atol=1e-8)
    for s in idx:
        n = ref["traj_n"][s]
        np.testing.assert_allclose(sol.t[s, :n], ref["traj"]["t"][s, :n], rtol=10 * rtol, atol=1e-9)
        np.testing.assert_allclose(sol.V[s, :n], ref["traj"]["V"][s, :n], rtol=rtol)
    return idx


def test_simulate_charge_and_discharge(P, lcoS, mS):
    """configs[4] physics at N = (10,10,10): a 1C charge (side reaction active: film grows, SOH drops) followed
    by a 1C discharge (inactive: ifelse(I_density > 0, ...), residuals.jl:546) on a randomised batch"""
    L = O.layout(mS)
    B = 32
    tho = util.oracle_theta_batch(B, first=6000)
    th = util.product_theta_from_oracle(lcoS, tho)
    util.set_theta_batch(lcoS, th)
    sol = P.simulate(lcoS, I=1, SOC=0, V_max=4.2)
    b = O.default_bounds("LCO", V_max=4.2)
    ref = O.simulate_batch(mS, tho, O.make_run("I", 1.0), O.default_opts(), b, SOC0=0.0, n_save_max=512, nthreads=8)
    idx = _compare(sol, ref)
    soh_ref = ref["state"]["Y"][:, L.SOH]
    s = sol.results[-1].summary
    np.testing.assert_allclose(s["aux_end"][idx], soh_ref[idx], rtol=1e-9)             # SOH of the final state
    assert np.all(s["aux_end"] < 1.0) and np.all(s["aux_end"] > 0.99)
    np.testing.assert_allclose(sol.Y[idx][:, L.film:L.film + 10], ref["state"]["Y"][idx][:, L.film:L.film + 10], rtol=1e-5)
    P.simulate_(sol, lcoS, I=-1, V_max=4.2)
    ref2 = O.simulate_batch(mS, tho, O.make_run("I", -1.0, new_run=False), O.default_opts(), b, state=ref["state"],
                            n_save_max=512, nthreads=8)
    s2 = sol.results[-1].summary
    same = (s2["n_steps"] == ref2["n_steps"]) & (s2["flag"] == ref2["flag"])
    print("discharge identical:", float(np.mean(same)))
    assert np.mean(same) >= 0.8
    np.testing.assert_allclose(s2["V_end"][same], ref2["V_end"][same], rtol=1e-6)
    np.testing.assert_allclose(s2["t_end"][same], ref2["t_end"][same], rtol=1e-5)


def test_soh_film_linear_invariant_on_device(P, lcoS, mS):
    """SOH - 1 = -(K rho_n/M_n) sum_i w_i film_i for every system, whatever steps the integrator took
    (tests/test_oracle_sei.py derives the weights independently): a known answer for the SEI rows"""
    from tests.test_oracle_sei import _K, soh_weights
    B = 64
    tho = util.oracle_theta_batch(B, first=900)
    th = util.product_theta_from_oracle(lcoS, tho)
    util.set_theta_batch(lcoS, th)
    sol = P.simulate(lcoS, I=1, SOC=0, V_max=4.2, reltol=1e-6, abstol=1e-6)
    names = O.theta_names()
    film = sol.Y[:, lcoS.ind["film"]]
    soh = sol.Y[:, lcoS.ind["SOH"]][:, 0]
    assert np.array_equal(soh, sol.results[-1].summary["aux_end"])
    assert film.min() > 0 and np.all(soh < 1)
    for s in range(B):
        thd = dict(zip(names, tho[s]))
        rhs = -_K(thd) * thd["rho_n"] / thd["M_n"] * np.dot(soh_weights(10, thd["l_n"]), film[s])
        np.testing.assert_allclose(soh[s] - 1.0, rhs, rtol=2e-5)


def test_discharge_without_film_resistance_is_the_model_without_aging(P, lcoS):
    """j_s = 0 on discharge (residuals.jl:546): with R_SEI = 0 the aging model IS the isothermal model -- the GPU's SEI
    family against the GPU's isothermal family at tight tolerance (their WRMS norms differ by sqrt(322/301))"""
    B = 16
    tho = util.oracle_theta_batch(B, first=940)
    tho[:, O.theta_names().index("R_SEI")] = 0.0
    iso = P.petlion("LCO")
    util.set_theta_batch(lcoS, util.product_theta_from_oracle(lcoS, tho))
    util.set_theta_batch(iso, util.product_theta_from_oracle(iso, tho))
    td = np.concatenate([np.arange(0.0, 3600.0, 60.0), [1e6]])
    a = P.simulate(lcoS, td, I=-1, SOC=1, reltol=1e-9, abstol=1e-9, n_save_max=0, outputs="all")
    b = P.simulate(iso, td, I=-1, SOC=1, reltol=1e-9, abstol=1e-9, n_save_max=0)
    sa, sb = a.results[-1].summary, b.results[-1].summary
    assert np.array_equal(sa["flag"], sb["flag"])
    # (a run that ends on V_min ends on the linear blend of its last step: h^2-accurate, see tests/test_gpu_tight.py)
    np.testing.assert_allclose(sa["t_end"], sb["t_end"], rtol=1e-4)
    fill = ~np.isnan(b.dense["V"])
    assert np.array_equal(fill, ~np.isnan(a.dense["V"]))
    np.testing.assert_allclose(a.dense["V"][fill], b.dense["V"][fill], rtol=1e-6)
    assert np.abs(a.Y[:, lcoS.ind["j_s"]]).max() < 1e-20 and np.abs(a.Y[:, lcoS.ind["film"]]).max() < 1e-24
    assert np.abs(sa["aux_end"] - 1.0).max() < 1e-14
