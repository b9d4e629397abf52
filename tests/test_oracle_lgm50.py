"""NMC_LGM50 / LiC6_LGM50 / system_LGM50_NMC_LiC6 (src/params.jl:514-849; the Chen et al. 2020 parameters of the LG M50
cell) in the CPU oracle.  Nothing in the reference executes this set (and its default aging = :stress throws), so it is
pinned to what the published cell implies and to itself: the 1C current density of a 5 Ah cell with 0.1027 m^2 of
electrode, the open-circuit window, the Jacobian of its own laws against central differences."""
import numpy as np

import oracle as O

CH = "NMC_LGM50"


def test_parameter_set_matches_the_published_cell():
    th = O.theta_defaults(CH)
    g = dict(zip(O.theta_names(), th))
    assert g["c_max_p"] == 63104.0 and g["c_max_n"] == 33133.0 and g["D_e"] == 8.794e-11 and g["t_plus"] == 0.2594
    i1c = O.calc_I1C(th)
    assert abs(i1c - 5.0 / 0.1027) / (5.0 / 0.1027) < 0.01          # A/m^2: the nominal 5 Ah over the electrode area
    b = O.default_bounds(CH)
    assert (b.V_min, b.V_max, b.T_max) == (2.5, 4.2, 55 + 273.15)
    m = O.make_model(CH)
    L = O.layout(m)
    # open-circuit voltage at the two ends of the stoichiometry window
    for soc, lo, hi in ((1.0, 4.1, 4.25), (0.0, 2.5, 3.3)):
        y0 = O.initial_guess(m, th, soc)
        V = y0[L.phi_s] - y0[L.phi_s + m.N_p + m.N_n - 1]
        assert lo < V < hi, (soc, V)


def test_discharge_and_charge_run_to_their_bounds():
    th = O.theta_defaults(CH)[None, :]
    for temperature in (False, True):
        m = O.make_model(CH, temperature=temperature)
        r = O.simulate_batch(m, th, O.make_run("I", -1.0), O.default_opts(), O.default_bounds(CH), SOC0=1.0)
        assert r["flag"][0] in (1, 3) and 3400 < r["t_end"][0] <= 3600.0 + 1e-6      # a full 1C discharge
        assert (r["T_end"][0] > 303.0) == temperature                               # it heats by ~9 K when allowed to
        r = O.simulate_batch(m, th, O.make_run("I", 1.0), O.default_opts(), O.default_bounds(CH), SOC0=0.0)
        assert r["flag"][0] == 2 and abs(r["V_end"][0] - 4.2) < 1e-9                 # charge ends on V_max = 4.2 V


def test_jacobian_against_central_differences():
    for temperature in (False, True):
        th = O.theta_defaults(CH)
        m = O.make_model(CH, temperature=temperature)
        L = O.layout(m)
        run = O.make_run("I", -1.0)
        Y = O.initial_guess(m, th, 0.55); Y[L.I] = -1.0
        it, Y, YP = O.newton_init(m, th, run, O.default_opts(), Y)
        assert it > 0
        # a state with gradients: a minute of discharge
        r = O.simulate_batch(m, th[None, :], O.make_run("I", -1.0, tf=60.0), O.default_opts(), O.default_bounds(CH), SOC0=0.55)
        Y, YP = r["state"]["Y"][0], r["state"]["YP"][0]
        gam = 0.21
        J = O.jacobian(m, th, run, 0.0, Y, YP, gam)
        cp, rv = O.jac_pattern(m, "I")
        worst = 0.0
        rng = np.random.default_rng(0)
        for c in rng.choice(L.N_tot, 60, replace=False):
            h = 1e-6 * max(abs(Y[c]), 1e-3)
            Yp_, Ym_ = Y.copy(), Y.copy(); Yp_[c] += h; Ym_[c] -= h
            YPp, YPm = YP.copy(), YP.copy(); YPp[c] += gam * h; YPm[c] -= gam * h
            d = (O.residual(m, th, run, 0.0, Yp_, YPp) - O.residual(m, th, run, 0.0, Ym_, YPm)) / (2 * h)
            col = np.zeros(L.N_tot); col[rv[cp[c]:cp[c + 1]]] = J[cp[c]:cp[c + 1]]
            worst = max(worst, np.abs(col - d).max() / max(np.abs(d).max(), 1e-30))
        assert worst < 1e-5, worst
