"""Ragged discretisations (sections of different sizes, odd node counts, short collectors): the lane mapping, the
twisted elimination (meeting node, chain lengths) and the collector folding must not depend on the 10/10/10 grid.
Same tolerances as tests/test_gpu_parity.py."""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu

GRIDS = [
    ("iso", dict(N_p=7, N_s=5, N_n=9), {}),
    ("iso", dict(N_p=12, N_s=3, N_n=17), {}),                       # 32 nodes: every lane owns one
    ("iso", dict(N_p=5, N_s=2, N_n=5), {}),
    ("wide", dict(N_p=13, N_s=9, N_n=17), {}),                      # 39 nodes: two warps, second one mostly idle
    ("wide", dict(N_p=20, N_s=24, N_n=20), {}),                     # 64 nodes
    ("thermal", dict(N_p=8, N_s=6, N_n=11, N_a=5, N_z=7), dict(temperature=True)),
    ("sei", dict(N_p=6, N_s=4, N_n=9), dict(aging=True)),
    ("wsei", dict(N_p=15, N_s=10, N_n=20), dict(aging=True)),
]


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


@pytest.mark.parametrize("family,grid,opt", GRIDS, ids=[f"{g[0]}-{'-'.join(str(v) for v in g[1].values())}" for g in GRIDS])
def test_ragged_grid_parity(P, family, grid, opt):
    p = P.petlion("LCO", **grid, **{k: ("SEI" if k == "aging" else v) for k, v in opt.items()})
    m = O.make_model("LCO", **grid, **opt)
    L = O.layout(m)
    assert p.N.tot == L.N_tot
    for method in ("I", "V"):
        cp, rv = O.jac_pattern(m, method)
        cp2, rv2 = p.jac_pattern(method)
        assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
    B = 6
    tho = util.oracle_theta_batch(B, first=500)
    th = util.product_theta_from_oracle(p, tho)
    util.set_theta_batch(p, th)
    # operator level: residual + Jacobian values at random physical states
    Y, YP = util.random_states(m, tho, seed=11)
    gam = np.full(B, 0.2)
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
            mask[L.T:L.T + grid["N_a"] + grid["N_p"] + grid["N_s"] + grid["N_n"] + grid["N_z"]] = False
        assert np.all(np.abs(res[s] - r_ref)[mask] <= tol * scale[mask] + 1e-300)
        rowmax = np.zeros(L.N_tot)
        np.maximum.at(rowmax, rv, np.abs(j_ref))
        assert np.all(np.abs(nz[s] - j_ref) <= 1e-6 * rowmax[rv] + 1e-300)
    # integrator level: a charge to a voltage bound, then the continuation
    sol = P.simulate(p, I=1.5, SOC=0.1, V_max=4.05)
    b = O.default_bounds("LCO", V_max=4.05)
    ref = O.simulate_batch(m, tho, O.make_run("I", 1.5), O.default_opts(), b, SOC0=0.1, n_save_max=512, nthreads=6)
    s = sol.results[-1].summary
    same = s["n_steps"] == ref["n_steps"]
    assert np.mean(same) >= 0.66, (s["n_steps"], ref["n_steps"])
    assert np.array_equal(s["flag"], ref["flag"]) and np.all(ref["flag"] == 2)
    np.testing.assert_allclose(s["t_end"][same], ref["t_end"][same], rtol=1e-6)
    np.testing.assert_allclose(s["t_end"], ref["t_end"], rtol=2e-3)
    np.testing.assert_allclose(s["SOC_end"], ref["SOC_end"], rtol=2e-3)
    P.simulate_(sol, p, 600, V="hold", V_max=4.05)
    ref2 = O.simulate_batch(m, tho, O.make_run("V", 0.0, tf=600, input_kind="hold", new_run=False), O.default_opts(), b,
                            state=ref["state"], n_save_max=512, nthreads=6)
    s2 = sol.results[-1].summary
    assert np.array_equal(s2["flag"], ref2["flag"])
    np.testing.assert_allclose(s2["I_end"], ref2["I_end"], rtol=5e-3, atol=1e-6)
    np.testing.assert_allclose(s2["SOC_end"], ref2["SOC_end"], rtol=2e-3)
