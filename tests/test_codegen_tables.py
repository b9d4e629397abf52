"""The generated particle operators the CUDA families are compiled with (petlion.jl_b200/csrc/laws_generated.cuh: namespaces
nr10 / nr12 / nr14 for the finite-difference scheme, sp10 for the spectral one) against the oracle's restatement of the
reference's residual, on the CPU: column k of kappa*MC is the c_s rows of the residual for c_s = e_k, the vector GJ (BJ e_surf in
the finite-difference scheme) is its response to j; EV diag(EL) EVI reproduces MC, EVIG = EVI GJ."""
import os
import re

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "petlion.jl_b200", "csrc", "laws_generated.cuh")).read()


def _ns(name):
    a = SRC.index(f"namespace {name} {{")
    return SRC[a:SRC.index(f"}}  // namespace {name}")]


def _mat(body, name):
    m = re.search(rf"double {name}\[NR\]\[NR\] = \{{(.*?)\n\}};", body, re.S)
    rows = re.findall(r"\{([^{}]*)\}", m.group(1))
    return np.array([[float(x) for x in r.split(",")] for r in rows])


def _vec(body, name):
    m = re.search(rf"double {name}\[NR\] = \{{([^}}]*)\}};", body)
    return np.array([float(x) for x in m.group(1).split(",")])


def _oracle_operator(nr, spectral):
    """(A, g): d(c_s rows of particle 0)/d(c_s of particle 0) / kappa and d(c_s rows)/dj * (-Rp), from the oracle's residual
    (which is linear in both)"""
    m = O.make_model("LCO", N_r_p=nr, N_r_n=nr, Fickian_method="spectral" if spectral else "finite_difference")
    L = O.layout(m)
    names = O.theta_names()
    th = O.theta_defaults("LCO")
    Ds, Rp = th[names.index("D_sp")], th[names.index("Rp_p")]
    kap = Ds / Rp**2
    Y0 = O.initial_guess(m, th, 0.5)
    Y0[L.c_s_p:L.c_s_p + nr] = 0.0
    Y0[L.j] = 0.0
    run = O.make_run("I", -1.0)
    zero = np.zeros_like(Y0)
    base = O.residual(m, th, run, 0.0, Y0, zero)[L.c_s_p:L.c_s_p + nr]
    assert np.all(base == 0.0)
    A = np.zeros((nr, nr))
    for k in range(nr):
        Y = Y0.copy(); Y[L.c_s_p + k] = 1.0
        A[:, k] = O.residual(m, th, run, 0.0, Y, zero)[L.c_s_p:L.c_s_p + nr] / kap
    Y = Y0.copy(); Y[L.j] = 1.0
    g = O.residual(m, th, run, 0.0, Y, zero)[L.c_s_p:L.c_s_p + nr] * (-Rp)      # rhs = kappa*(MC c + GJ d1bc), kappa*d1bc = -j/Rp
    return A, g


@pytest.mark.parametrize("name,nr,spectral", [("nr10", 10, False), ("nr12", 12, False), ("nr14", 14, False), ("sp10", 10, True)])
def test_generated_operator_equals_the_oracles(name, nr, spectral):
    body = _ns(name)
    assert int(re.search(r"constexpr int NR = (\d+);", body).group(1)) == nr
    MC, EV, EVI, EL = _mat(body, "MC"), _mat(body, "EV"), _mat(body, "EVI"), _vec(body, "EL")
    A, g = _oracle_operator(nr, spectral)
    scale = np.abs(A).max()
    np.testing.assert_allclose(MC, A, rtol=0, atol=2e-12 * scale)
    if spectral:
        GJ, EVIG = _vec(body, "GJ"), _vec(body, "EVIG")
        np.testing.assert_allclose(GJ, g, rtol=1e-11, atol=1e-11 * np.abs(g).max())
        np.testing.assert_allclose(EVIG, EVI @ GJ, rtol=1e-10, atol=1e-12 * np.abs(EVIG).max())
        assert int(re.search(r"mc_mask\(int\) \{ return (0x[0-9a-f]+)u", body).group(1), 16) == (1 << nr) - 1
    else:
        BJ = float(re.search(r"constexpr double BJ = ([^;]+);", body).group(1))
        want = np.zeros(nr); want[-1] = BJ
        np.testing.assert_allclose(g, want, rtol=1e-12, atol=1e-12 * BJ)
        masks = [int(x, 16) for x in re.findall(r"r == \d+ \? (0x[0-9a-f]+)u", body)]
        assert len(masks) == nr
        for r in range(nr):
            assert masks[r] == sum(1 << c for c in range(nr) if A[r, c] != 0.0), r
    # the eigen-decomposition the structured solve runs in: real, and it reproduces the operator
    np.testing.assert_allclose(EV @ np.diag(EL) @ EVI, MC, rtol=0, atol=1e-10 * scale)
    np.testing.assert_allclose(EV @ EVI, np.eye(nr), atol=1e-12)
    # conservation (MC 1 = 0 up to the rounding of the stencil's rows) and a stable spectrum
    assert abs(EL[-1]) < 1e-15 * abs(EL[0]) and np.all(EL[:-1] < 0.0)
