"""The oracle's restatement of Fickian_method = :spectral (residuals_c_s_avg!, residuals.jl:181-235; "BETA" in the reference,
executed by none of its examples or tests: unpinned).  What the scheme itself implies is checked here:
  * Chebyshev collocation differentiates polynomials of degree < N_r exactly: for c_s(r) = a + b r^2 + d r^4 with the surface
    flux that matches it, every radial row of dc_s/dt equals D_s/Rp^2 (6b + 20 d r^2) -- including the centre row (L'Hopital);
  * a 1C discharge agrees with the finite-difference scheme to a few mV (two discretisations of the same particle);
  * the Jacobian pattern: a dense 10x10 block and a full j column per particle."""
import numpy as np
import pytest

import oracle as O


def _radii(nr):
    x = np.cos(np.pi * np.arange(nr) / (nr - 1))
    return ((x + 1.0) / 2.0)[::-1]          # stored centre -> surface


@pytest.mark.parametrize("nr", [8, 10, 12, 16])
def test_polynomial_profile_has_the_exact_rate(nr):
    m = O.make_model("LCO", N_r_p=nr, N_r_n=nr, Fickian_method="spectral")
    L = O.layout(m)
    names = O.theta_names()
    th = O.theta_defaults("LCO")
    Y = O.initial_guess(m, th, 0.5)
    r = _radii(nr)
    cases = (("p", L.c_s_p, m.N_p, th[names.index("D_sp")], th[names.index("Rp_p")], 30000.0, -700.0, 150.0),
             ("n", L.c_s_n, m.N_n, th[names.index("D_sn")], th[names.index("Rp_n")], 12000.0, 450.0, -90.0))
    for el, off, n, Ds, Rp, a, b, d in cases:
        for e in range(n):
            f = 1.0 + 0.05 * e
            Y[off + e * nr: off + (e + 1) * nr] = a + f * (b * r**2 + d * r**4)
            Y[L.j + (e if el == "p" else m.N_p + e)] = -f * (2.0 * b + 4.0 * d) * Ds / Rp      # dc/dr(1) = -j Rp / D_s
    res = O.residual(m, th, O.make_run("I", -1.0), 0.0, Y, np.zeros_like(Y))
    for el, off, n, Ds, Rp, a, b, d in cases:
        for e in range(n):
            f = 1.0 + 0.05 * e
            want = Ds / Rp**2 * f * (6.0 * b + 20.0 * d * r**2)
            got = res[off + e * nr: off + (e + 1) * nr]
            np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-9 * np.abs(want).max(), err_msg=f"{el} particle {e}, N_r={nr}")


def test_pattern_and_discharge_next_to_finite_differences():
    ms = O.make_model("LCO", Fickian_method="spectral"); mf = O.make_model("LCO")
    cps, _ = O.jac_pattern(ms, "I"); cpf, _ = O.jac_pattern(mf, "I")
    assert cps[-1] - cpf[-1] == 20 * (18 + 9)
    th = O.theta_defaults("LCO")[None, :]
    rs = O.simulate_batch(ms, th, O.make_run("I", -1.0), O.default_opts(), O.default_bounds("LCO"), SOC0=1.0)
    rf = O.simulate_batch(mf, th, O.make_run("I", -1.0), O.default_opts(), O.default_bounds("LCO"), SOC0=1.0)
    assert rs["flag"][0] == rf["flag"][0] == 3 and abs(rs["t_end"][0] - 3600.0) < 1e-6 and abs(rf["t_end"][0] - 3600.0) < 1e-6
    assert abs(rs["V_end"][0] - rf["V_end"][0]) < 5e-3
