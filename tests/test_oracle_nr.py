"""The oracle's particle stencil for N_r != 10 (residuals_c_s_avg!, residuals.jl:128-180; derivative matrices
numerical_tools.jl:8-87).  The reference executes only N_r = 10 (every example), so other N_r are unpinned against it; what the
scheme itself implies is checked here: the 9-point / 5-point stencils are exact on low-degree polynomials, so for
c_s(r) = a + b r^2 with the surface flux that matches it (dc/dr|_{r=1} = 2b = -j Rp / D_s) every radial row of
dc_s/dt equals the spherical Laplacian D_s/Rp^2 * 6b -- for any N_r >= 9 the matrices can be built for."""
import numpy as np
import pytest

import oracle as O


@pytest.mark.parametrize("nr", [9, 10, 11, 12, 14, 16, 20])
def test_quadratic_profile_has_uniform_rate(nr):
    m = O.make_model("LCO", N_r_p=nr, N_r_n=nr)
    L = O.layout(m)
    assert L.N_tot == 2 * 30 + (nr + 2) * 20 + 1
    names = O.theta_names()
    th = O.theta_defaults("LCO")
    Y = O.initial_guess(m, th, 0.5)
    r = np.linspace(0.0, 1.0, nr)
    for el, off, n, Ds, Rp, a, b in (("p", L.c_s_p, m.N_p, th[names.index("D_sp")], th[names.index("Rp_p")], 30000.0, -700.0),
                                     ("n", L.c_s_n, m.N_n, th[names.index("D_sn")], th[names.index("Rp_n")], 12000.0, 450.0)):
        for e in range(n):
            bb = b * (1.0 + 0.05 * e)
            Y[off + e * nr: off + (e + 1) * nr] = a + bb * r * r
            Y[L.j + (e if el == "p" else m.N_p + e)] = -2.0 * bb * Ds / Rp
    res = O.residual(m, th, O.make_run("I", -1.0), 0.0, Y, np.zeros_like(Y))
    for el, off, n, Ds, Rp, b in (("p", L.c_s_p, m.N_p, th[names.index("D_sp")], th[names.index("Rp_p")], -700.0),
                                  ("n", L.c_s_n, m.N_n, th[names.index("D_sn")], th[names.index("Rp_n")], 450.0)):
        for e in range(n):
            want = Ds / Rp**2 * 6.0 * b * (1.0 + 0.05 * e)
            got = res[off + e * nr: off + (e + 1) * nr]
            np.testing.assert_allclose(got, want, rtol=2e-9, err_msg=f"{el} particle {e}, N_r={nr}")


@pytest.mark.parametrize("nr", [12, 14])
def test_particle_block_pattern_grows_by_nine_per_node(nr):
    """the Jacobian pattern of a particle: 82 entries at N_r = 10 and nine more per added radial node (interior rows carry the
    9-point first-derivative stencil); the j column adds one entry per particle"""
    m10 = O.make_model("LCO"); m = O.make_model("LCO", N_r_p=nr, N_r_n=nr)
    cp10, _ = O.jac_pattern(m10, "I"); cp, _ = O.jac_pattern(m, "I")
    assert cp[-1] - cp10[-1] == 20 * 9 * (nr - 10)
