"""Known answers for the aging = :SEI rows.  The reference ships no executed SEI example, so these rows cannot be
pinned to printed digits; what CAN be pinned are the closed forms the equations imply
(src/physics_equations/residuals.jl:260-297, 519-552; src/external.jl:469-522):
  * film' = -j_s M_n/rho_n, row by row;
  * SOH'  = K * trapz(extrapolate_section(j_s)),  K = F a_n / (3600 I1C): for a uniform j_s = c this is K c l_n, and
    for any profile it is a fixed linear functional (weights derived here independently, in numpy);
  * therefore SOH - 1 + (K rho_n/M_n) sum_i w_i film_i is a LINEAR INVARIANT of the DAE (film(0) = 0, SOH(0) = 1),
    which a BDF integrator preserves to the Newton tolerance for ANY step sequence;
  * j_s = 0 on discharge (I_density > 0 is false, :546): film = 0, SOH = 1 stay exact, and with R_SEI = 0 the voltage
    is the one of the model without aging (compared at tight tolerance: the WRMS norms of the two models differ by sqrt(N)).
"""
import numpy as np

import oracle as O

F = 96485.3321233


def soh_weights(N, l_n):
    """d trapz(extrapolate_section(y, :n)) / d y_i, derived independently of oracle.c (external.jl:469-522)."""
    x = np.concatenate([[0.0], np.linspace(1 / (2 * N), 1 - 1 / (2 * N), N), [1.0]])

    def extrap_x_0(xx, yy):       # second-order polynomial through three points, evaluated at x = 0
        c = np.polyfit(xx, yy, 2)
        return np.polyval(c, 0.0)
    w = np.zeros(N)
    for i in range(N):
        y = np.zeros(N); y[i] = 1.0
        y0 = extrap_x_0(x[1:4], y[:3])
        # the right end uses the SAME abscissae x[2:4] with the reversed values (external.jl:512): distance from the end
        yN = extrap_x_0(x[1:4], y[::-1][:3])
        yr = np.concatenate([[y0], y, [yN]])
        xs = x * l_n
        w[i] = np.sum(0.5 * np.diff(xs) * (yr[1:] + yr[:-1]))
    return w


def _K(th):
    eps_sn = 1.0 - (th["eps_fn"] + th["eps_n"])
    return F * (3 * eps_sn / th["Rp_n"]) / (3600 * O.calc_I1C(np.array(list(th.values()))))


def _state(ms, th, soc=0.4, cur=1.0):
    L = O.layout(ms)
    y0 = O.initial_guess(ms, th, soc); y0[L.I] = cur
    it, y, yp = O.newton_init(ms, th, O.make_run("I", cur), O.default_opts(), y0)
    assert it > 0
    return L, y, yp


def test_film_and_soh_rows_closed_form():
    ms = O.make_model("LCO", aging=True)
    th = O.theta_defaults("LCO"); thd = dict(zip(O.theta_names(), th))
    L, y, yp = _state(ms, th)
    rng = np.random.default_rng(7)
    run = O.make_run("I", 1.0)
    K = _K(thd)
    for trial in range(4):
        yy, ypp = y.copy(), yp.copy()
        js = -1e-8 * (1 + rng.uniform(size=ms.N_n)) if trial else np.full(ms.N_n, -1.3e-8)
        yy[L.j_s:L.j_s + ms.N_n] = js
        ypp[L.film:L.film + ms.N_n] = rng.normal(size=ms.N_n) * 1e-15
        ypp[L.SOH] = rng.normal() * 1e-9
        res = O.residual(ms, th, run, 0.0, yy, ypp)
        # film rows
        np.testing.assert_allclose(res[L.film:L.film + ms.N_n] + ypp[L.film:L.film + ms.N_n],
                                   -js * thd["M_n"] / thd["rho_n"], rtol=1e-14)
        # SOH row
        rhs = res[L.SOH] + ypp[L.SOH]
        if trial == 0:
            np.testing.assert_allclose(rhs, K * js[0] * thd["l_n"], rtol=1e-12)     # uniform profile: K c l_n
        np.testing.assert_allclose(rhs, K * np.dot(soh_weights(ms.N_n, thd["l_n"]), js), rtol=1e-11)


def test_side_reaction_rate_law():
    """j_s row on charge: j_s + |i0 (I/I1C)^w / F * exp(-0.5 F eta_s / (R T))| with eta_s built from j + j_s and the
    film resistance (residuals.jl:519-552); zero on discharge."""
    ms = O.make_model("LCO", aging=True)
    th = O.theta_defaults("LCO"); thd = dict(zip(O.theta_names(), th))
    L, y, yp = _state(ms, th, cur=2.0)
    y[L.film:L.film + ms.N_n] = np.linspace(1e-9, 5e-9, ms.N_n)
    R = 8.31446261815324
    n0 = ms.N_p
    for cur in (2.0, 0.5, -1.0):
        y[L.I] = cur
        res = O.residual(ms, th, O.make_run("I", cur), 0.0, y, yp)
        j = y[L.j + n0:L.j + n0 + ms.N_n]; js = y[L.j_s:L.j_s + ms.N_n]
        ps = y[L.phi_s + n0:L.phi_s + n0 + ms.N_n]; pe = y[L.phi_e + ms.N_p + ms.N_s:L.phi_e + L.Nx]
        Rf = thd["R_SEI"] + y[L.film:L.film + ms.N_n] / thd["k_n_aging"]
        eta_s = ps - pe - thd["Uref_s"] - F * (j + js) * Rf
        calc = -np.abs(thd["i_0_jside"] * cur ** thd["w"] / F * (-np.exp(-0.5 * F / (R * thd["T0"]) * eta_s))) if cur > 0 else 0.0
        np.testing.assert_allclose(res[L.j_s:L.j_s + ms.N_n], js - calc, rtol=1e-12, atol=1e-24)


def test_soh_film_linear_invariant():
    """SOH - 1 = -(K rho_n/M_n) sum_i w_i film_i along any trajectory, whatever steps the integrator takes."""
    ms = O.make_model("LCO", aging=True)
    th = O.theta_defaults("LCO"); thd = dict(zip(O.theta_names(), th))
    L = O.layout(ms)
    w = soh_weights(ms.N_n, thd["l_n"])
    for tol in (1e-3, 1e-6):
        o = O.default_opts(reltol=tol, abstol=min(tol, 1e-6), reltol_init=tol, abstol_init=min(tol, 1e-6))
        r = O.simulate_batch(ms, th[None, :], O.make_run("I", 1.0), o, O.default_bounds("LCO", V_max=4.2), SOC0=0.0)
        Y = r["state"]["Y"][0]
        film = Y[L.film:L.film + ms.N_n]
        assert film.min() > 0 and Y[L.SOH] < 1.0
        lhs = Y[L.SOH] - 1.0
        rhs = -_K(thd) * thd["rho_n"] / thd["M_n"] * np.dot(w, film)
        # both sides are ~1e-4; the integrator's Newton iteration leaves rows unconverged at ~reltol of their scale
        np.testing.assert_allclose(lhs, rhs, rtol=20 * tol)


def test_discharge_has_no_side_reaction_and_matches_the_model_without_aging():
    ms, mi = O.make_model("LCO", aging=True), O.make_model("LCO")
    th = O.theta_defaults("LCO")[None, :].copy()
    # the aging model carries the film resistance R_SEI + film/k_n_aging in eta even when nothing grows
    # (auxiliary_states_and_coefficients.jl:272-300): without R_SEI the two models are the same physics
    th[0, O.theta_names().index("R_SEI")] = 0.0
    Ls = O.layout(ms)
    o = O.default_opts(reltol=1e-9, abstol=1e-9, reltol_init=1e-9, abstol_init=1e-9)
    td = np.arange(0.0, 3600.0, 60.0)
    a = O.simulate_batch(ms, th, O.make_run("I", -1.0), o, O.default_bounds("LCO"), SOC0=1.0, dense_t=td, dense_Y=True)
    b = O.simulate_batch(mi, th, O.make_run("I", -1.0), o, O.default_bounds("LCO"), SOC0=1.0, dense_t=td)
    Y = a["dense"]["Y"][0]
    # (zero up to the round-off of the linear solves: j is ~1e-5 mol/m^2/s, a charge grows ~1e-11 m of film)
    assert np.abs(Y[:, Ls.j_s:Ls.j_s + ms.N_n]).max() < 1e-20 and np.abs(Y[:, Ls.film:Ls.film + ms.N_n]).max() < 1e-24
    assert np.abs(Y[:, Ls.SOH] - 1.0).max() < 1e-14
    assert a["flag"][0] == b["flag"][0] == 3
    np.testing.assert_allclose(a["dense"]["V"], b["dense"]["V"], rtol=1e-6)
    np.testing.assert_allclose(a["t_end"], b["t_end"], rtol=1e-9)
