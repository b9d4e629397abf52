"""temperature = true together with aging = :SEI (params.jl:119-174 allows the combination): N = 372 for N = (10,10,10).
The side-reaction rate feels the node temperature, the heat source F a j_total (T dU/dT + eta) feels j_s and the film.
UNPINNED against the reference (nothing in it executes aging = :SEI); GPU against the oracle's restatement."""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


@pytest.fixture(scope="module", params=["10-10-10", "20-20-20"])
def fam(request, P):
    """one warp per system (N = 372) and, on 60 x-nodes, two warps per system (N = 722)"""
    grid = {} if request.param == "10-10-10" else dict(N_p=20, N_s=20, N_n=20)
    return P.petlion("LCO", temperature=True, aging="SEI", **grid), O.make_model("LCO", temperature=True, aging=True, **grid)


def test_sizes_and_patterns(fam):
    p, m = fam
    L = O.layout(m)
    assert p.N.tot == L.N_tot == (372 if m.N_p == 10 else 722) and p.N.diff == L.N_diff
    for method in ("I", "V", "P", "dT", "η_p"):
        cp, rv = O.jac_pattern(m, "eta_p" if method == "η_p" else method)
        cp2, rv2 = p.jac_pattern(method)
        assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2), method


def _states(m, tho, cur, soc0, t_mid):
    b = O.default_bounds("LCO", V_max=4.3)
    r = O.simulate_batch(m, tho, O.make_run("I", cur, tf=t_mid), O.default_opts(), b, SOC0=soc0, nthreads=8)
    assert np.all(r["flag"] == 0)
    return r["state"]["Y"], r["state"]["YP"]


@pytest.mark.parametrize("cur,soc0,method,value", [(2.0, 0.2, "I", 2.0), (-1.0, 0.9, "V", 3.8), (1.5, 0.3, "P", 150.0), (2.0, 0.3, "dT", 0.0)])
def test_resjac_parity(fam, cur, soc0, method, value):
    p, m = fam
    B = 8
    N = O.layout(m).N_tot
    tho = util.oracle_theta_batch(B, first=10)
    th = util.product_theta_from_oracle(p, tho)
    Y, YP = _states(m, tho, cur, soc0, 500.0)
    gam = np.random.default_rng(5).uniform(0.01, 50.0, size=B)
    res, nz = p.resjac(Y, YP, gam, method=method, value=value, theta=th)
    run = O.make_run(method, value)
    cp, rv = O.jac_pattern(m, method)
    cols = np.repeat(np.arange(N), np.diff(cp))
    for s in range(B):
        r_ref = O.residual(m, tho[s], run, 0.0, Y[s], YP[s])
        j_ref = O.jacobian(m, tho[s], run, 0.0, Y[s], YP[s], gam[s])
        scale = np.zeros(N)
        np.maximum.at(scale, rv, np.abs(j_ref) * np.maximum(np.abs(Y[s][cols]), 1e-12))
        scale = np.maximum(scale, np.abs(r_ref))
        er = np.abs(res[s] - r_ref) / (scale + 1e-300)
        assert er.max() < 1e-9, (s, int(er.argmax()), res[s][er.argmax()], r_ref[er.argmax()])
        rowmax = np.zeros(N); np.maximum.at(rowmax, rv, np.abs(j_ref))
        ej = np.abs(nz[s] - j_ref) / rowmax[rv]
        k = int(ej.argmax())
        assert ej.max() < 1e-9, (s, int(rv[k]), int(cols[k]), nz[s][k], j_ref[k])


@pytest.mark.parametrize("method,value", [("I", 2.0), ("V", 4.0), ("dT", 0.0)])
def test_linear_solve_equals_dense(fam, method, value):
    p, m = fam
    B = 4
    tho = util.oracle_theta_batch(B, first=40)
    th = util.product_theta_from_oracle(p, tho)
    Y, YP = _states(m, tho, 2.0, 0.2, 500.0)       # charging: the side reaction and its I column are live
    gam = np.array([50.0, 1.0, 0.05, 0.01])
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
        print("thsei solve", method, s, rr, rr_ref)
        # (the dT row -- the weighted mean of all Y'[T] -- against rows of size 1e+5: same bound as the thermal family's
        #  own dT solve test, tests/test_gpu_thermal.py)
        assert rr < (max(100 * rr_ref, 1e-7) if method == "dT" else 50 * rr_ref + 1e-11), (s, rr, rr_ref)


def test_newton_init_parity(fam):
    p, m = fam
    L = O.layout(m)
    B = 6
    tho = util.oracle_theta_batch(B)
    th = util.product_theta_from_oracle(p, tho)
    soc = np.linspace(0.1, 0.9, B)
    cur = np.where(np.arange(B) % 2 == 0, -1.0, 2.0)
    Y0 = p.initial_guess(soc, theta=th)
    for s in range(B):
        np.testing.assert_allclose(Y0[s], O.initial_guess(m, tho[s], soc[s]), rtol=1e-13, atol=0)
    Y0[:, L.I] = cur
    st, Y, YP = p.newton_init(Y0, method="I", value=cur, theta=th)
    for s in range(B):
        it, y, yp = O.newton_init(m, tho[s], O.make_run("I", cur[s]), O.default_opts(), Y0[s])
        assert st[s] == it
        np.testing.assert_allclose(Y[s], y, rtol=1e-9, atol=1e-14)


def test_fast_charge_protocol_with_aging(P, fam):
    """4C charge -> dT = :hold -> V = :hold, then a 1C discharge: SOH falls during the three charging segments only,
    the cell heats up, every segment against the oracle"""
    p, m = fam
    L = O.layout(m)
    B = 16
    tho = util.oracle_theta_batch(B, first=2000)
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    segs = [("I", "value", 4.0, 1e6, {"V_max": 4.1, "T_max": 310.0}), ("dT", "hold", 0.0, 300.0, {"V_max": 4.1}),
            ("V", "hold", 0.0, 600.0, {"V_max": 4.15}), ("I", "value", -1.0, 1200.0, {})]
    W = dict(cathode="LCO", temperature=True, aging=True, soc0=0.0, segs=segs,
             grid={} if m.N_p == 10 else dict(N_p=20, N_s=20, N_n=20))
    ref = util.oracle_protocol(W, tho, O.default_opts())
    sol, _ = util.gpu_protocol(P, p, W)
    same = np.ones(B, dtype=bool)
    soh = []
    for k in range(len(segs)):
        s, r = sol.results[k].summary, ref[k]
        for c in ("flag", "n_steps", "n_res", "n_jac", "n_netf", "n_ncfn"):
            same &= s[c] == r[c]
        print("segment", k, "identical so far", same.mean(), s["n_steps"][:6], r["n_steps"][:6], s["flag"][:6], r["flag"][:6])
        assert (s["flag"] >= 0).all() and (r["flag"] >= 0).all()
        assert same.mean() >= 0.5
        np.testing.assert_allclose(s["V_end"][same], r["V_end"][same], rtol=5e-6)
        np.testing.assert_allclose(s["T_end"][same], r["T_end"][same], rtol=1e-6)    # (0.3 mK; the dT row is ill-conditioned, see above)
        np.testing.assert_allclose(s["aux_end"][same], r["state"]["Y"][same][:, L.SOH], rtol=1e-7)
        np.testing.assert_allclose(s["V_end"], r["V_end"], rtol=5e-3)
        np.testing.assert_allclose(s["T_end"], r["T_end"], rtol=5e-3)
        np.testing.assert_allclose(s["aux_end"], r["state"]["Y"][:, L.SOH], rtol=1e-5)
        soh.append(s["aux_end"].copy())
    assert np.all(soh[0] < 1.0) and np.all(soh[1] < soh[0]) and np.all(soh[2] < soh[1])
    np.testing.assert_array_equal(soh[3], soh[2])          # j_s == 0 on discharge (residuals.jl:546)
    assert np.all(sol.results[0].summary["T_end"] > 299.0)


def test_tight_tolerance_whole_trajectory(P):
    """north_star's tolerance on the new family: at reltol = abstol = 1e-7 (the thermal families' floor, see
    tests/test_gpu_tight.py) every row of V / T / SOC and of the SOH up to 60 s before the exit agrees to 1e-6 for every
    system -- a 2C charge with the side reaction running, whatever steps either side takes"""
    B = 128
    p = P.petlion("LCO", temperature=True, aging="SEI")
    m = O.make_model("LCO", temperature=True, aging=True)
    L = O.layout(m)
    tho = util.oracle_theta_batch(B, first=5000)
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    tol = 1e-7
    td = np.arange(7.0, 1700.0, 20.0)
    sol = P.simulate(p, 1700.0, I=2, SOC=0.0, V_max=4.2, dense_t=td, n_save_max=0, reltol=tol, abstol=tol, maxiters=100000)
    o = O.default_opts(reltol=tol, abstol=tol, reltol_init=tol, abstol_init=tol, maxiters=100000)
    ref = O.simulate_batch(m, tho, O.make_run("I", 2.0, tf=1700.0), o, O.default_bounds("LCO", V_max=4.2), SOC0=0.0, nthreads=16,
                           dense_t=td, dense_Y=True)
    s = sol.results[0].summary
    assert (s["flag"] >= 0).all() and (ref["flag"] >= 0).all()
    before_exit = td[None, :] <= np.minimum(s["t_end"], ref["t_end"])[:, None] - 60.0
    for key in ("V", "I", "SOC", "T"):
        g, r = sol.dense[key], ref["dense"][key]
        both = ~np.isnan(g) & ~np.isnan(r) & before_exit
        assert both.sum() > 20 * B
        err = np.abs(g - r)[both] / (np.maximum(np.abs(r[both]), 1e-3) if key != "SOC" else 1.0)
        assert err.max() <= 1e-6, (key, float(err.max()))
    # the capacity fade at the end: SOH of the final state
    np.testing.assert_allclose(s["aux_end"], ref["state"]["Y"][:, L.SOH], rtol=1e-6)
    assert np.all(s["aux_end"] < 0.9999)
