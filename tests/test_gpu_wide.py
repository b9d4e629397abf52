"""GPU parity tests of the wide families: grids with 33..64 x-nodes, two warps per system.
BASELINE.json configs[4]: LCO, aging = :SEI, N = (20,20,20) -> 642 DAEs; the same grid without aging (601); and with
temperature = true (681: N_a = N_z = 10)."""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu

GRID = dict(N_p=20, N_s=20, N_n=20)


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


@pytest.fixture(scope="module", params=["plain", "sei", "thermal"])
def fam(request, P):
    aging = request.param == "sei"
    thermal = request.param == "thermal"
    p = P.petlion("LCO", aging="SEI" if aging else False, temperature=thermal, **GRID)
    m = O.make_model("LCO", aging=aging, temperature=thermal, **GRID)
    return p, m, aging


def test_sizes_and_pattern(fam):
    p, m, aging = fam
    L = O.layout(m)
    assert p.N.tot == L.N_tot == (642 if aging else (681 if m.temperature else 601))
    assert p.N.diff == L.N_diff
    for method in ("I", "V", "P"):
        cp, rv = O.jac_pattern(m, method)
        cp2, rv2 = p.jac_pattern(method)
        assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2), method


def _states(m, tho, cur, soc0, t_mid):
    b = O.default_bounds("LCO", V_max=4.3)
    r = O.simulate_batch(m, tho, O.make_run("I", cur, tf=t_mid), O.default_opts(), b, SOC0=soc0, nthreads=8)
    assert np.all(r["flag"] == 0)
    return r["state"]["Y"], r["state"]["YP"]


@pytest.mark.parametrize("cur,soc0,method,value", [(1.0, 0.2, "I", 1.0), (-1.0, 0.9, "V", 3.8), (1.5, 0.3, "P", 150.0)])
def test_resjac_parity(fam, cur, soc0, method, value):
    p, m, aging = fam
    B = 6
    L = O.layout(m); N = L.N_tot
    tho = util.oracle_theta_batch(B, first=10)
    th = util.product_theta_from_oracle(p, tho)
    Y, YP = _states(m, tho, cur, soc0, 600.0)
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


def test_linear_solve_equals_dense(fam):
    p, m, aging = fam
    B = 4
    tho = util.oracle_theta_batch(B, first=40)
    th = util.product_theta_from_oracle(p, tho)
    Y, YP = _states(m, tho, 1.0, 0.2, 600.0)
    gam = np.array([50.0, 1.0, 0.05, 0.01])
    run = O.make_run("I", 1.0)
    cp, rv = O.jac_pattern(m, "I")
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
    x, st = p.linear_solve(Y, YP, gam, rhs, method="I", value=1.0, theta=th)
    for s in range(B):
        xr = np.linalg.solve(Js[s], rhs[s])
        rr = np.linalg.norm(Js[s] @ x[s] - rhs[s]) / np.linalg.norm(rhs[s])
        rr_ref = np.linalg.norm(Js[s] @ xr - rhs[s]) / np.linalg.norm(rhs[s])
        print("wide solve", aging, s, rr, rr_ref)
        assert rr < 20 * rr_ref + 1e-11, (s, rr, rr_ref)


def test_newton_init_parity(fam):
    p, m, aging = fam
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
    opts = O.default_opts()
    for s in range(B):
        it, y, yp = O.newton_init(m, tho[s], O.make_run("I", cur[s]), opts, Y0[s])
        assert st[s] == it
        np.testing.assert_allclose(Y[s], y, rtol=1e-9, atol=1e-14)


def test_simulate_parity(P, fam):
    """configs[4] at test size: randomised batch; 1C charge (side reaction active with aging) and 1C discharge"""
    p, m, aging = fam
    B = 16
    tho = util.oracle_theta_batch(B, first=8000)
    th = util.product_theta_from_oracle(p, tho)
    util.set_theta_batch(p, th)
    b = O.default_bounds("LCO", V_max=4.2)
    td = np.arange(7.0, 3700.0, 45.0)
    for cur, soc0 in ((1.0, 0.0), (-1.0, 1.0)):
        sol = P.simulate(p, I=cur, SOC=soc0, V_max=4.2, dense_t=td)
        ref = O.simulate_batch(m, tho, O.make_run("I", cur), O.default_opts(), b, SOC0=soc0, n_save_max=512, nthreads=8, dense_t=td)
        s = sol.results[-1].summary
        same = np.ones(B, dtype=bool)          # the same decisions: every counter agrees
        for c in ("flag", "n_steps", "n_res", "n_jac", "n_netf", "n_ncfn"):
            same &= s[c] == ref[c]
        print("wide", aging, cur, "identical", float(np.mean(same)), s["n_steps"][:6], ref["n_steps"][:6], s["flag"][:6], ref["flag"][:6])
        assert np.mean(same) >= 0.75
        # thermal: the conduction rows carry ~1e-5 K/s of cancellation noise (tests/test_gpu_thermal.py).  It enters the error
        # norms, hence the step-size factors: with equal counters the step TIMES still drift apart by up to ~1e-3 relative, so
        # rows saved at each side's own step times are not comparable there; compare on a common grid (both sides' dense
        # output) -- which is also what bounds the systems that took different decisions
        np.testing.assert_allclose(s["V_end"][same], ref["V_end"][same], rtol=2e-5 if m.temperature else 1e-6)
        np.testing.assert_allclose(s["V_end"], ref["V_end"], rtol=5e-3)        # the others: same answer at the integrator's tolerance
        np.testing.assert_allclose(s["t_end"][same], ref["t_end"][same], rtol=1e-5)
        g, r = sol.dense["V"], ref["dense"]["V"]
        both = ~np.isnan(g) & ~np.isnan(r) & (td[None, :] <= np.minimum(s["t_end"], ref["t_end"])[:, None] - 30.0)
        assert both[same].sum() > 30 * same.sum()
        err = np.where(both, np.abs(g - r) / np.maximum(np.abs(r), 1e-3), 0.0)
        assert err[same].max() <= (2e-5 if m.temperature else 1e-6), float(err[same].max())
        # systems that took different decisions: 5 reltol on the common grid, up to 90 s before the earlier exit (at reltol 1e-3 the
        # last steps of a discharge are tens of seconds long and V falls 15 mV/s in the knee: a 1 s shift of the knee is 0.5 %)
        knee = td[None, :] > np.minimum(s["t_end"], ref["t_end"])[:, None] - 90.0
        worst = np.unravel_index(np.argmax(np.where(knee, 0.0, err)), err.shape)
        assert np.where(knee, 0.0, err).max() <= 5e-3, (float(err[worst]), int(worst[0]), float(td[worst[1]]), float(s["t_end"][worst[0]]), float(ref["t_end"][worst[0]]))
        assert err.max() <= 2e-2, float(err.max())
        if not m.temperature:
            for k in np.where(same)[0]:
                n = ref["traj_n"][k]
                np.testing.assert_allclose(sol.V[k, :n], ref["traj"]["V"][k, :n], rtol=1e-6)
        if aging and cur > 0:
            L = O.layout(m)
            np.testing.assert_allclose(s["aux_end"][same], ref["state"]["Y"][same][:, L.SOH], rtol=1e-9)
        if m.temperature:
            np.testing.assert_allclose(s["T_end"][same], ref["T_end"][same], rtol=1e-6)      # 0.3 mK (conduction-row noise, see above)
            assert np.all(s["T_end"] > 298.2)


def test_wide_thermal_protocol_on_a_ragged_grid(P):
    """temperature = true on 36 x-nodes with unequal sections and collectors: 2C charge -> dT = :hold -> V = :hold
    (the README's CC-CT-CV shape), every segment against the oracle"""
    grid = dict(N_p=12, N_s=8, N_n=16, N_a=7, N_z=9)
    p = P.petlion("LCO", temperature=True, **grid)
    m = O.make_model("LCO", temperature=True, **grid)
    assert p.N.tot == O.layout(m).N_tot == 2 * 36 + 12 * 28 + 1 + 36 + 16
    B = 8
    tho = util.oracle_theta_batch(B, first=300)
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    segs = [("I", "value", 3.0, 400.0, {}), ("dT", "hold", 0.0, 300.0, {}), ("V", "hold", 0.0, 300.0, {})]
    W = dict(cathode="LCO", temperature=True, grid=grid, soc0=0.1, segs=segs)
    ref = util.oracle_protocol(W, tho, O.default_opts())
    sol, _ = util.gpu_protocol(P, p, W)
    for k in range(3):
        s, r = sol.results[k].summary, ref[k]
        same = np.ones(B, dtype=bool)
        for c in ("flag", "n_steps", "n_res", "n_jac", "n_netf", "n_ncfn"):
            same &= s[c] == r[c]
        print("segment", k, "identical", same.mean(), s["n_steps"], r["n_steps"])
        assert same.mean() >= 0.6 and (s["flag"] >= 0).all()
        np.testing.assert_allclose(s["V_end"][same], r["V_end"][same], rtol=1e-6)
        np.testing.assert_allclose(s["T_end"][same], r["T_end"][same], rtol=1e-6)
        np.testing.assert_allclose(s["I_end"][same], r["I_end"][same], rtol=1e-5, atol=1e-8)
        # decision flips: still the same answer at the integrator's tolerance
        np.testing.assert_allclose(s["V_end"], r["V_end"], rtol=5e-3)
        np.testing.assert_allclose(s["T_end"], r["T_end"], rtol=5e-3)
