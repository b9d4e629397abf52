"""rxn_MHC (Marcus-Hush-Chidsey kinetics, src/physics_equations/custom_functions.jl:233-298) as the reaction rate law
of either electrode: `petlion(LCO; rxn_p=rxn_MHC, rxn_n=rxn_MHC)` (src/params.jl:163-170).

UNPINNED against the reference (nothing in it executes rxn_MHC): the oracle restates the function from its source,
the GPU carries hand-derived partial derivatives; the tests compare the two.

The reference's parameter sets give lambda_MHC_p = lambda_MHC_n = 6.26e-20 (params.jl:16, 67 -- a reorganisation energy in
Joules where the function expects it in units of kT): every erf of the function is then -1 and the law degenerates
to 2 k (c_e0 c_s sigma(eta_f) - c_e c_max sigma(-eta_f)) sqrt((1 - theta)/c_e0).  Both that default and a physical
value (lambda = 15, 8) are run.
"""
import numpy as np
import pytest

import oracle as O
from tests import util

pytestmark = pytest.mark.gpu
RX = dict(rxn_p="rxn_MHC", rxn_n="rxn_MHC")


@pytest.fixture(scope="module")
def P():
    import petlion_b200
    return petlion_b200


def _theta(B, lam, first=0):
    tho = util.oracle_theta_batch(B, first=first)
    if lam is not None:
        names = O.theta_names()
        tho[:, names.index("lambda_MHC_p")] = lam
        tho[:, names.index("lambda_MHC_n")] = 0.8 * lam
    return tho


def test_keys_and_defaults(P):
    p = P.petlion("LCO", **RX)
    keys = list(p.θ_keys)
    # code-point order of the reference: 'M' < 'a', so the two new keys sit right before λ_a's place (here: after θ_min_p)
    assert keys.index("λ_MHC_n") + 1 == keys.index("λ_MHC_p") and len(keys) == 37
    assert p.θ["λ_MHC_p"] == 6.26e-20 and p.θ["λ_MHC_n"] == 6.26e-20
    assert len(P.petlion("LCO", rxn_p="rxn_MHC").θ_keys) == 36
    assert P.model_key(p) != P.model_key(P.petlion("LCO"))
    with pytest.raises(Exception, match="lambda_MHC"):
        P.petlion("NMC", **RX)          # the NMC set has no λ_MHC_*: KeyError in the reference
    with pytest.raises(ValueError):
        P.petlion("LCO", rxn_p="rxn_Tafel")


@pytest.mark.parametrize("family", ["iso", "thermal", "sei", "mixed"])
@pytest.mark.parametrize("lam", [None, 15.0])
def test_resjac_parity(P, family, lam):
    kw = dict(temperature=family == "thermal", aging="SEI" if family == "sei" else False)
    rx = dict(rxn_p="rxn_BV", rxn_n="rxn_MHC") if family == "mixed" else RX
    p = P.petlion("LCO", **kw, **rx)
    m = O.make_model("LCO", temperature=kw["temperature"], aging=bool(kw["aging"]), **rx)
    N = O.layout(m).N_tot
    B = 24
    tho = _theta(B, lam)
    tho[B // 2:, O.theta_names().index("T0")] = 305.0
    th = util.product_theta_from_oracle(p, tho)
    Y, YP = util.random_states(m, tho, seed=11)
    gam = np.random.default_rng(5).uniform(0.01, 50.0, size=B)
    res, nz = p.resjac(Y, YP, gam, method="I", value=1.0, theta=th)
    run = O.make_run("I", 1.0)
    cp, rv = O.jac_pattern(m, "I")
    cp2, rv2 = p.jac_pattern("I")
    assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
    for s in range(B):
        r_ref = O.residual(m, tho[s], run, 0.0, Y[s], YP[s])
        j_ref = O.jacobian(m, tho[s], run, 0.0, Y[s], YP[s], gam[s])
        scale = np.zeros(N)
        for c in range(N):
            k = slice(cp[c], cp[c + 1])
            np.maximum.at(scale, rv[k], np.abs(j_ref[k]) * max(abs(Y[s][c]), 1e-12))
        scale = np.maximum(scale, np.abs(r_ref))
        assert np.all(np.abs(res[s] - r_ref) <= 1e-10 * scale + 1e-300), (s, int(np.argmax(np.abs(res[s] - r_ref) / scale)))
        rowmax = np.zeros(N)
        np.maximum.at(rowmax, rv, np.abs(j_ref))
        err = np.abs(nz[s] - j_ref) / rowmax[rv]
        assert err.max() < 1e-9, (s, int(np.argmax(err)), err.max())


@pytest.mark.parametrize("lam", [None, 15.0, 8.0])
def test_discharge_batch_takes_the_oracle_steps(P, lam):
    B = 96
    p = P.petlion("LCO", **RX)
    m = O.make_model("LCO", **RX)
    tho = _theta(B, lam, first=500)
    util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
    td = np.arange(7.0, 3700.0, 60.0)
    sol = P.simulate(p, 1e6, I=-1, SOC=1.0, dense_t=td, n_save_max=0)
    ref = O.simulate_batch(m, tho, O.make_run("I", -1.0, tf=1e6), O.default_opts(), O.default_bounds(), nthreads=16, dense_t=td)
    s = sol.results[0].summary
    same = np.ones(B, dtype=bool)
    for k in ("flag", "n_steps", "n_res", "n_jac", "n_netf", "n_ncfn"):
        same &= s[k] == ref[k]
    # (with the reference's default lambda a good part of the batch ends in IDA_CONV_FAIL near the end of the
    # discharge -- on both sides, the same systems)
    assert same.mean() >= (0.9 if lam is None else 0.95), (same.mean(), np.unique(s["flag"], return_counts=True), np.unique(ref["flag"], return_counts=True))
    ok = same & (ref["flag"] >= 0)
    assert lam is None or ok.sum() > 0.9 * B          # (default lambda: every system of this batch ends in -2, on both sides)
    assert np.abs(s["V_end"] - ref["V_end"])[ok].max(initial=0.0) <= 1e-6 * 4.0
    g, o = sol.dense["V"][same], ref["dense"]["V"][same]
    both = ~np.isnan(g) & ~np.isnan(o)
    assert both.sum() > 40 * same.sum() and np.abs(g[both] - o[both]).max() <= 1e-6 * 4.0
    # the law matters: the voltage differs from the Butler-Volmer model's by millivolts
    q = P.petlion("LCO")
    util.set_theta_batch(q, util.product_theta_from_oracle(q, tho))
    bv = P.simulate(q, 1e6, I=-1, SOC=1.0, dense_t=td, n_save_max=0)
    d = np.abs(bv.dense["V"] - sol.dense["V"])
    assert np.nanmax(d) > 1e-3


def test_tight_tolerance_whole_trajectory(P):
    """reltol = abstol = 1e-9, physical lambda: every row of V / I / SOC up to the exit within 1e-6, thermal model included"""
    for temperature, tol, B in ((False, 1e-9, 128), (True, 1e-7, 64)):
        p = P.petlion("LCO", temperature=temperature, **RX)
        m = O.make_model("LCO", temperature=temperature, **RX)
        tho = _theta(B, 15.0, first=900)
        util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
        td = np.arange(7.0, 3400.0, 30.0)
        sol = P.simulate(p, 3400.0, I=-1, SOC=1.0, dense_t=td, n_save_max=0, reltol=tol, abstol=tol, maxiters=100000)
        o = O.default_opts(reltol=tol, abstol=tol, reltol_init=tol, abstol_init=tol, maxiters=100000)
        ref = O.simulate_batch(m, tho, O.make_run("I", -1.0, tf=3400.0), o, O.default_bounds(), nthreads=16, dense_t=td)
        s = sol.results[0].summary
        assert (s["flag"] >= 0).all() and (ref["flag"] >= 0).all()
        for key in ("V", "I", "SOC") + (("T",) if temperature else ()):
            g, r = sol.dense[key], ref["dense"][key]
            both = ~np.isnan(g) & ~np.isnan(r) & (td[None, :] <= np.minimum(s["t_end"], ref["t_end"])[:, None] - 60.0)
            err = np.abs(g - r)[both] / (np.maximum(np.abs(r[both]), 1e-3) if key != "SOC" else 1.0)
            # (thermal: run at 1e-7, see tests/test_gpu_tight.py; two BDF integrators whose error tests flip on round-off are
            #  each ~10 tol from the exact solution at the end of a discharge: 20 tol between them)
            assert err.max() <= (2e-6 if temperature else 1e-6), (temperature, key, float(err.max()))
