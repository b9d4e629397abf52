"""NMC_LGM50 (src/params.jl:514-849, Chen et al. 2020) on the GPU: `petlion(NMC_LGM50; aging=false)`, isothermal and with
temperature = true (the set's default).  Its own OCVs and electrolyte laws are a third instantiation of the iso / thermal
families.  UNPINNED against the reference (nothing in it executes this set); GPU against the oracle's restatement."""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu
CH = "NMC_LGM50"


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


@pytest.fixture(scope="module", params=[False, True], ids=["isothermal", "thermal"])
def fam(request, P):
    return P.petlion(CH, temperature=request.param), O.make_model(CH, temperature=request.param)


def test_defaults_and_keys(P, fam):
    p, m = fam
    assert P.petlion(CH).numerics.temperature is True           # system_LGM50_NMC_LiC6 defaults to temperature = true
    g = dict(zip(O.theta_names(), O.theta_defaults(CH)))
    names = {"T₀": "T0", "c_e₀": "c_e0", "t₊": "t_plus"}
    for k in p.θ_keys:
        a = names.get(k, k.replace("θ", "theta").replace("λ", "lambda").replace("ρ", "rho").replace("σ", "sigma").replace("ϵ", "eps"))
        assert p.θ[k] == g[a], (k, p.θ[k], g[a])
    assert "D_e" in p.θ_keys and "D_p" not in p.θ_keys and len(p.θ_keys) == (54 if m.temperature else 33)
    assert p.bounds.V_min == 2.5 and p.bounds.V_max == 4.2 and p.bounds.T_max == 55 + 273.15
    assert abs(p.I1C()[0] - O.calc_I1C(O.theta_defaults(CH))) < 1e-12
    assert P.model_key(p).startswith("NMC_LGM50_LiC6_LGM50/")
    for method in ("I", "V", "P"):
        cp, rv = O.jac_pattern(m, method)
        cp2, rv2 = p.jac_pattern(method)
        assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2), method
    with pytest.raises(RuntimeError, match="LCO parameter set"):
        P.petlion(CH, aging="SEI")
    # (33..64 x-nodes: the two-warp instantiations of this chemistry, tests/test_gpu_matrix.py)


def _states(m, tho, cur, soc0, t_mid):
    r = O.simulate_batch(m, tho, O.make_run("I", cur, tf=t_mid), O.default_opts(), O.default_bounds(CH), SOC0=soc0, nthreads=8)
    assert np.all(r["flag"] == 0)
    return r["state"]["Y"], r["state"]["YP"]


@pytest.mark.parametrize("cur,soc0,method,value", [(-1.0, 0.9, "I", -1.0), (1.0, 0.2, "V", 3.9), (1.5, 0.3, "P", 150.0)])
def test_resjac_parity(fam, cur, soc0, method, value):
    p, m = fam
    B = 12
    N = O.layout(m).N_tot
    tho = util.oracle_theta_batch(B, cathode=CH, first=10)
    tho[B // 2:, O.theta_names().index("T0")] = 305.0           # Arrhenius branches of D_s and k
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


def test_newton_init_parity(fam):
    p, m = fam
    L = O.layout(m)
    B = 6
    tho = util.oracle_theta_batch(B, cathode=CH)
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


def test_simulate_parity(P, fam):
    """randomised batch: 1C discharge to V_min / SOC_min, 1C charge to 4.2 V, then V = :hold; common grid through both sides'
    dense output (see tests/test_gpu_wide.py for why not at each side's own step times in the thermal family)"""
    p, m = fam
    B = 32
    tho = util.oracle_theta_batch(B, cathode=CH, first=3000)
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    td = np.arange(7.0, 5000.0, 45.0)
    W = dict(cathode=CH, temperature=bool(m.temperature), soc0=1.0,
             segs=[("I", "value", -1.0, 1e6, {}), ("I", "value", 1.0, 1e6, {}), ("V", "hold", 0.0, 600.0, {})])
    ref = util.oracle_protocol(W, tho, O.default_opts(), dense_t=td)
    sol, dense = util.gpu_protocol(P, p, W, dense_t=td)
    same = np.ones(B, dtype=bool)
    for k in range(3):
        s, r = sol.results[k].summary, ref[k]
        for c in ("flag", "n_steps", "n_res", "n_jac", "n_netf", "n_ncfn"):
            same &= s[c] == r[c]
        print("LGM50", bool(m.temperature), "segment", k, "identical so far", same.mean(), s["n_steps"][:6], r["n_steps"][:6], s["flag"][:6], r["flag"][:6])
        assert (s["flag"] >= 0).all() and (r["flag"] >= 0).all()
        assert same.mean() >= 0.6
        tolV = 2e-5 if m.temperature else 1e-6
        np.testing.assert_allclose(s["V_end"][same], r["V_end"][same], rtol=tolV)
        np.testing.assert_allclose(s["V_end"], r["V_end"], rtol=5e-3)
        np.testing.assert_allclose(s["T_end"], r["T_end"], rtol=5e-3)
        g, o = dense[k]["V"], r["dense"]["V"]
        both = ~np.isnan(g) & ~np.isnan(o) & (td[None, :] <= np.minimum(s["t_end"], r["t_end"])[:, None] - 30.0)
        err = np.where(both, np.abs(g - o) / np.maximum(np.abs(o), 1e-3), 0.0)
        assert err[same].max(initial=0.0) <= tolV and err.max() <= 5e-3, (k, float(err[same].max(initial=0.0)), float(err.max()))
    if m.temperature:
        assert np.all(sol.results[0].summary["T_end"] > 300.0)
