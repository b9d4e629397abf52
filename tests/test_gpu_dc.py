"""The concentration-rate inputs dc_s_p_max .. dc_e_min (src/physics_equations/input_methods.jl:190-245) on the GPU: a
continuation run whose control row is  val - Y'[ind],  ind = the arg-max / arg-min surface (or electrolyte) concentration of
the state the run starts from.  Built for the isothermal families without aging (a sibling build of each, chosen per run).
UNPINNED against the reference (nothing in it executes them): GPU against the oracle, and against the closed form the
control equation implies -- the chosen state moves at exactly the requested rate."""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu
KINDS = ("dc_s_p_max", "dc_s_p_min", "dc_s_n_max", "dc_s_n_min", "dc_e_max", "dc_e_min")


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


def _index(m, L, Y, kind):
    if kind.startswith("dc_e"):
        v = Y[:, L.c_e:L.c_e + L.Nx]
        return L.c_e + (v.argmax(axis=1) if kind.endswith("max") else v.argmin(axis=1))
    n = kind[5] == "n"
    base, cnt, Nr = (L.c_s_n, m.N_n, m.N_r_n) if n else (L.c_s_p, m.N_p, m.N_r_p)
    v = Y[:, base + Nr - 1:base + cnt * Nr:Nr]
    return base + Nr - 1 + Nr * (v.argmax(axis=1) if kind.endswith("max") else v.argmin(axis=1))


@pytest.mark.parametrize("grid", [{}, dict(N_p=20, N_s=20, N_n=20), dict(N_p=7, N_s=5, N_n=9)])
@pytest.mark.parametrize("cathode", ["LCO", "NMC"])
def test_rate_inputs_against_the_oracle(P, grid, cathode):
    if cathode == "NMC" and grid:
        pytest.skip("one grid is enough for the second parameter set")
    B = 12
    p = P.petlion(cathode, **grid)
    m = O.make_model(cathode, **grid)
    L = O.layout(m)
    tho = util.oracle_theta_batch(B, cathode=cathode, first=700)
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    opts = O.default_opts()
    b = O.default_bounds(cathode)
    r0 = O.simulate_batch(m, tho, O.make_run("I", -1.0, tf=600.0), opts, b, nthreads=8)
    for kind in KINDS:
        sol = P.simulate(p, 600.0, I=-1, SOC=1.0)
        Y0, YP0 = sol.Y.copy(), sol.YP.copy() if hasattr(sol, "YP") and sol.YP is not None else None
        idx = _index(m, L, Y0, kind)
        np.testing.assert_array_equal(idx, _index(m, L, r0["state"]["Y"], kind))
        rate = 0.5 * r0["state"]["YP"][np.arange(B), idx]
        for inp, kindname, want in ((rate, "value", rate * 300.0), ("hold", "hold", np.zeros(B))):
            s2 = P.simulate(p, 600.0, I=-1, SOC=1.0)
            P.simulate_(s2, p, 300.0, **{kind: inp})
            ref = O.simulate_batch(m, tho, O.make_run(kind, 0.0, tf=300.0, input_kind=kindname, new_run=False), opts, b,
                                   values=None if kindname == "hold" else rate, state=r0["state"], nthreads=8)
            s = s2.results[-1].summary
            assert (s["flag"] == ref["flag"]).all() and (s["flag"] >= 0).all(), (kind, kindname, s["flag"], ref["flag"])
            same = np.ones(B, dtype=bool)
            for c in ("n_steps", "n_res", "n_jac", "n_netf", "n_ncfn"):
                same &= s[c] == ref[c]
            assert same.mean() >= 0.7, (kind, kindname, s["n_steps"], ref["n_steps"])
            np.testing.assert_allclose(s["V_end"][same], ref["V_end"][same], rtol=1e-6)
            np.testing.assert_allclose(s["I_end"][same], ref["I_end"][same], rtol=1e-5, atol=1e-8)
            np.testing.assert_allclose(s["V_end"], ref["V_end"], rtol=5e-3)
            np.testing.assert_allclose(s["I_end"], ref["I_end"], rtol=5e-2, atol=5e-3)
            # the closed form: the chosen state moved by rate * time (at the integrator's tolerance)
            got = s2.Y[np.arange(B), idx] - Y0[np.arange(B), idx]
            np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-3 * np.abs(Y0[np.arange(B), idx]).max() * 1e-3)


def test_tight_tolerance_follows_the_closed_form(P):
    p = P.petlion("LCO")
    m = O.make_model("LCO"); L = O.layout(m)
    tol = dict(reltol=1e-9, abstol=1e-9)
    for kind in ("dc_s_n_min", "dc_e_max"):
        sol = P.simulate(p, 600.0, I=-1, SOC=1.0, **tol)
        Y0 = sol.Y.copy()
        idx = int(_index(m, L, Y0, kind)[0])
        P.simulate_(sol, p, 200.0, **{kind: -0.25}, **tol)
        assert abs((sol.Y[0, idx] - Y0[0, idx]) - (-0.25 * 200.0)) < 1e-5
        P.simulate_(sol, p, 200.0, **{kind: "hold"}, **tol)            # a NEW arg-min is taken from the current state
        assert sol.results[-1].summary["flag"][0] == 0


def test_refusals(P):
    p = P.petlion("LCO")
    with pytest.raises(ValueError, match="previous solution"):
        P.simulate(p, 10.0, dc_e_max=0.0)
    sol = P.simulate(p, 10.0, I=-1, SOC=1.0)
    with pytest.raises(ValueError, match="number or :hold"):
        P.simulate_(sol, p, 10.0, dc_e_max="rest")
    for kw in (dict(temperature=True), dict(aging="SEI")):
        q = P.petlion("LCO", **kw)
        s = P.simulate(q, 10.0, I=-1, SOC=1.0)
        with pytest.raises(RuntimeError, match="isothermal models without aging"):
            P.simulate_(s, q, 10.0, dc_s_n_max=0.0)
