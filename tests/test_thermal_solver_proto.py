"""CPU test: the lane-level prototype of the thermal variant's structured solver (the algorithm the CUDA
code in csrc/plb_device.cuh implements: eigen-basis particle solve, current-collector chains, twisted
4x4 block-Thomas with the two-node-wide T-row couplings absorbed exactly, applied-current border)
reproduces a dense LAPACK solve of the oracle's Jacobian for all three control modes."""
import numpy as np
import pytest

import oracle as O
from tests import proto_thermal_solver as P


def _state():
    m = O.make_model("LCO", temperature=True); th = O.theta_defaults("LCO")
    b = O.default_bounds("LCO", T_max=40 + 273.15, V_max=4.1)
    r = O.simulate_batch(m, th, O.make_run("I", 4.0, tf=120.0), O.default_opts(), b, SOC0=0.0)
    return m, th, r["state"]["Y"][0], r["state"]["YP"][0]


def test_inv4_formula():
    rng = np.random.default_rng(3)
    a = rng.normal(size=(7, 4, 4)) + 3 * np.eye(4)
    np.testing.assert_allclose(P.inv4(a), np.linalg.inv(a), rtol=1e-11, atol=1e-13)


def test_layout_matches_oracle():
    g = P.Geo(); Lo = O.layout(O.make_model("LCO", temperature=True))
    assert (g.N, g.off_T, g.off_j, g.off_pe, g.off_ps, g.off_I) == (Lo.N_tot, Lo.T, Lo.j, Lo.phi_e, Lo.phi_s, Lo.I)


@pytest.mark.parametrize("method,val", [("I", 4.0), ("V", 4.0), ("P", 300.0)])
@pytest.mark.parametrize("cj", [50.0, 0.5, 0.01])
def test_structured_solve_equals_dense(method, val, cj):
    m, th, Y, YP = _state()
    g = P.Geo()
    run = O.make_run(method, val)
    cp, rv = O.jac_pattern(m, method)
    nz = O.jacobian(m, th, run, 0.0, Y, YP, cj)
    J = np.zeros((g.N, g.N))
    for c in range(g.N):
        J[rv[cp[c]:cp[c + 1]], c] = nz[cp[c]:cp[c + 1]]
    Fa = P.factor(g, P.lane_jac_from_dense(g, J, cj), cj)
    rhs = np.random.default_rng(1).normal(size=g.N) * np.abs(J).max(axis=1) * 1e-3
    x = P.solve(g, Fa, rhs)
    xr = np.linalg.solve(J, rhs)
    # backward-stable: same residual as LAPACK (the matrices have condition numbers 1e15..1e18)
    rr = np.linalg.norm(J @ x - rhs) / np.linalg.norm(rhs)
    rr_ref = np.linalg.norm(J @ xr - rhs) / np.linalg.norm(rhs)
    assert rr < 5 * rr_ref + 1e-13
    assert np.max(np.abs(x - xr)) / np.max(np.abs(xr)) < 1e-6
