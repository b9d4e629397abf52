"""The rest of the option matrix (second sitting of round 2): rxn_MHC on the two-warp grids and with temperature and aging
together, NMC_LGM50 on the two-warp grids -- further instantiations of the same templates (csrc/plb_variant_{widemhc,wseimhc,
wthmhc,thseimhc,wthseimhc,widelgm,wthlgm}.cu), each checked against the oracle like its 32-node sibling: Jacobian pattern,
residual / Jacobian values, the structured linear solve against dense LAPACK, the Newton initialisation, and a charge to a
voltage bound + CV hold.  Tolerances as in tests/test_gpu_wide.py / tests/test_gpu_mhc.py."""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu
G = dict(N_p=20, N_s=20, N_n=20)
RX = dict(rxn_p="rxn_MHC", rxn_n="rxn_MHC")
CASES = {
    "widemhc": ("LCO", dict(**G, **RX), 15.0),
    "wseimhc": ("LCO", dict(aging=True, **G, **RX), 15.0),
    "wthmhc": ("LCO", dict(temperature=True, **G, **RX), 15.0),
    "thseimhc": ("LCO", dict(temperature=True, aging=True, **RX), 15.0),
    "wthseimhc": ("LCO", dict(temperature=True, aging=True, **G, **RX), 12.0),
    "widelgm": ("NMC_LGM50", dict(temperature=False, **G), None),
    "wthlgm": ("NMC_LGM50", dict(temperature=True, N_p=12, N_s=10, N_n=14), None),   # 36 nodes, ragged
}


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


@pytest.fixture(scope="module", params=list(CASES), ids=list(CASES))
def fam(request, P):
    cathode, kw, lam = CASES[request.param]
    p = P.petlion(cathode, **{k: ("SEI" if k == "aging" and v else v) for k, v in kw.items()})
    m = O.make_model(cathode, **kw)
    return request.param, cathode, p, m, lam


def _theta(cathode, B, lam, first=0):
    tho = util.oracle_theta_batch(B, cathode=cathode, first=first)
    if lam is not None:
        names = O.theta_names()
        tho[:, names.index("lambda_MHC_p")] = lam
        tho[:, names.index("lambda_MHC_n")] = 0.8 * lam
    return tho


def test_pattern_resjac_solve_newton(fam):
    name, cathode, p, m, lam = fam
    L = O.layout(m)
    N = L.N_tot
    assert p.N.tot == N
    methods = ("I", "V", "P") + (("dT",) if m.temperature else ())
    for method in methods:
        cp, rv = O.jac_pattern(m, method)
        cp2, rv2 = p.jac_pattern(method)
        assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2), method
    B = 4
    tho = _theta(cathode, B, lam, first=60)
    th = util.product_theta_from_oracle(p, tho)
    Y, YP = util.random_states(m, tho, seed=7)
    gam = np.array([50.0, 1.0, 0.05, 0.01])
    for method, value in (("I", 1.0), ("V", 4.0)):
        res, nz = p.resjac(Y, YP, gam, method=method, value=value, theta=th)
        run = O.make_run(method, value)
        cp, rv = O.jac_pattern(m, method)
        cols = np.repeat(np.arange(N), np.diff(cp))
        Js = []
        for s in range(B):
            r_ref = O.residual(m, tho[s], run, 0.0, Y[s], YP[s])
            j_ref = O.jacobian(m, tho[s], run, 0.0, Y[s], YP[s], gam[s])
            scale = np.zeros(N)
            np.maximum.at(scale, rv, np.abs(j_ref) * np.maximum(np.abs(Y[s][cols]), 1e-12))
            scale = np.maximum(scale, np.abs(r_ref))
            mask = np.ones(N, bool)
            if m.temperature:      # T rows: the reference's A_tot*T form carries ~1e-5 K/s of cancellation noise
                mask[L.T:L.T + (L.film if m.aging else L.j) - L.T] = False
            er = (np.abs(res[s] - r_ref) / (scale + 1e-300))[mask]
            assert er.max() < 1e-9, (method, s, float(er.max()))
            rowmax = np.zeros(N); np.maximum.at(rowmax, rv, np.abs(j_ref))
            ej = np.abs(nz[s] - j_ref) / rowmax[rv]
            assert ej.max() < 1e-6, (method, s, float(ej.max()))
            J = np.zeros((N, N)); J[rv, cols] = j_ref
            Js.append(J)
        rng = np.random.default_rng(2)
        rhs = np.stack([rng.normal(size=N) * np.abs(J).max(axis=1) * 1e-3 for J in Js])
        x, st = p.linear_solve(Y, YP, gam, rhs, method=method, value=value, theta=th)
        for s in range(B):
            xr = np.linalg.solve(Js[s], rhs[s])
            rr = np.linalg.norm(Js[s] @ x[s] - rhs[s]) / np.linalg.norm(rhs[s])
            rr_ref = np.linalg.norm(Js[s] @ xr - rhs[s]) / np.linalg.norm(rhs[s])
            assert rr < 50 * rr_ref + 1e-11, (method, s, rr, rr_ref)
    soc = np.linspace(0.15, 0.85, B)
    cur = np.where(np.arange(B) % 2 == 0, -1.0, 1.5)
    Y0 = p.initial_guess(soc, theta=th)
    Y0[:, L.I] = cur
    st, Yn, YPn = p.newton_init(Y0, method="I", value=cur, theta=th)
    for s in range(B):
        it, y, yp = O.newton_init(m, tho[s], O.make_run("I", cur[s]), O.default_opts(), Y0[s])
        assert st[s] == it
        np.testing.assert_allclose(Yn[s], y, rtol=1e-9, atol=1e-12)


def test_charge_and_hold(P, fam):
    name, cathode, p, m, lam = fam
    B = 12
    tho = _theta(cathode, B, lam, first=4000)
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    td = np.arange(7.0, 4000.0, 45.0)
    vmax = 4.1
    W = dict(cathode=cathode, temperature=bool(m.temperature), aging=bool(m.aging), soc0=0.1,
             grid=dict(N_p=m.N_p, N_s=m.N_s, N_n=m.N_n), rx={k: v for k, v in CASES[name][1].items() if k.startswith("rxn_")},
             segs=[("I", "value", 1.0, 1e6, {"V_max": vmax}), ("V", "hold", 0.0, 600.0, {"V_max": vmax})])
    # (util.oracle_protocol builds its model from cathode / temperature / aging / grid: pass the rate law through the model here)
    out, state = [], None
    for k, (method, kind, value, tf, bo) in enumerate(W["segs"]):
        b = O.default_bounds(cathode, **bo)
        r = O.simulate_batch(m, tho, O.make_run(method, value, tf=tf, input_kind=kind, new_run=(k == 0)), O.default_opts(), b,
                             SOC0=W["soc0"], state=state, nthreads=8, dense_t=td)
        state = r["state"]; out.append(r)
    sol, dense = util.gpu_protocol(P, p, W, dense_t=td)
    same = np.ones(B, dtype=bool)
    for k in range(2):
        s, r = sol.results[k].summary, out[k]
        for c in ("flag", "n_steps", "n_res", "n_jac", "n_netf", "n_ncfn"):
            same &= s[c] == r[c]
        print(name, "segment", k, "identical so far", same.mean(), s["n_steps"][:6], r["n_steps"][:6], s["flag"][:6], r["flag"][:6])
        assert np.array_equal(s["flag"] >= 0, r["flag"] >= 0)
        ok = (s["flag"] >= 0) & (r["flag"] >= 0)
        assert ok.mean() >= 0.75 and (same | ~ok).mean() >= 0.5
        tolV = 2e-5 if m.temperature else 1e-6
        np.testing.assert_allclose(s["V_end"][same & ok], r["V_end"][same & ok], rtol=tolV)
        np.testing.assert_allclose(s["V_end"][ok], r["V_end"][ok], rtol=5e-3)
        g, o = dense[k]["V"], r["dense"]["V"]
        both = ~np.isnan(g) & ~np.isnan(o) & ok[:, None] & (td[None, :] <= np.minimum(s["t_end"], r["t_end"])[:, None] - 30.0)
        err = np.where(both, np.abs(g - o) / np.maximum(np.abs(o), 1e-3), 0.0)
        assert err[same].max(initial=0.0) <= tolV and err.max() <= 5e-3, (k, float(err[same].max(initial=0.0)), float(err.max()))
