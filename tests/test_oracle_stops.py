"""CPU tests of the oracle's stop predicates with their t_frac back-interpolation (src/checks.jl:141-224,
model_evaluation.jl:369-382): c_s_n_max, c_e_min, eta_plating_min, dfilm_max.  The final state is the linear
blend Y_prev + t_frac (Y - Y_prev), and each bounded quantity is linear in Y, so it must sit ON its bound."""
import numpy as np

import oracle as O


def _run(model, I, soc, opts=None, **bo):
    th = O.theta_defaults("LCO")[None, :]
    return O.simulate_batch(model, th, O.make_run("I", I), opts or O.default_opts(), O.default_bounds("LCO", **bo),
                            SOC0=soc, n_save_max=600), dict(zip(O.theta_names(), th[0]))


def test_stop_c_s_n_max():
    m = O.make_model("LCO"); L = O.layout(m)
    r, th = _run(m, 2.0, 0.0, c_s_n_max=0.5)
    Y = r["state"]["Y"][0]
    assert r["flag"][0] == 6
    surf = Y[L.c_s_n + m.N_r_n - 1:L.c_s_n + m.N_n * m.N_r_n:m.N_r_n]
    assert abs(surf.max() / th["c_max_n"] - 0.5) < 1e-12
    # not during a discharge (checks.jl:148: only for I > 0)
    r2, _ = _run(m, -1.0, 1.0, c_s_n_max=0.5)
    assert r2["flag"][0] == 3
    # the last saved row is the interpolated end point
    n = r["traj_n"][0]
    assert r["traj"]["t"][0, n - 1] == r["t_end"][0] and r["traj"]["t"][0, n - 2] < r["t_end"][0]


def test_stop_c_e_min():
    m = O.make_model("LCO"); L = O.layout(m)
    r, _ = _run(m, 2.0, 0.0, c_e_min=800.0)
    assert r["flag"][0] == 9
    assert abs(r["state"]["Y"][0][L.c_e:L.c_e + L.Nx].min() - 800.0) < 1e-9
    # also on discharge (no current-sign test, checks.jl:170)
    r2, _ = _run(m, -2.0, 1.0, c_e_min=800.0)
    assert r2["flag"][0] == 9
    # (the node that holds the minimum moves during this step: the blend of the two minima is a lower bound)
    assert 800.0 - 1e-9 <= r2["state"]["Y"][0][L.c_e:L.c_e + L.Nx].min() < 801.0


def test_stop_eta_plating_min():
    m = O.make_model("LCO"); L = O.layout(m)
    for bound in (0.1, 0.05, 0.02):
        r, _ = _run(m, 4.0, 0.0, eta_plating_min=bound)
        Y = r["state"]["Y"][0]
        assert r["flag"][0] == 11
        assert abs((Y[L.phi_s + m.N_p] - Y[L.phi_e + m.N_p + m.N_s]) - bound) < 1e-12
    # times are ordered: a lower bound trips later
    t = [_run(m, 4.0, 0.0, eta_plating_min=b)[0]["t_end"][0] for b in (0.1, 0.05, 0.02)]
    assert t[0] < t[1] < t[2]


def test_stop_dfilm_max():
    ms = O.make_model("LCO", aging=True); L = O.layout(ms)
    free, th = _run(ms, 1.0, 0.0, V_max=4.2)
    assert free["flag"][0] in (2, 4)
    r, _ = _run(ms, 1.0, 0.0, V_max=4.2, dfilm_max=3e-15)
    assert r["flag"][0] == 10 and r["t_end"][0] < free["t_end"][0]
    # film' = -j_s M_n/rho_n (residuals.jl:260-276): the growth rate of the blended end state is on the bound to
    # the accuracy of IDA's derivative interpolant (the predicate looks at Y', the blend acts on Y)
    Y = r["state"]["Y"][0]
    rate = (-Y[L.j_s:L.j_s + ms.N_n] * th["M_n"] / th["rho_n"]).max()
    assert abs(rate - 3e-15) < 2e-2 * 3e-15
    # a discharge has no side reaction (residuals.jl:546) and never trips it
    r3, _ = _run(ms, -1.0, 1.0, dfilm_max=1e-18)
    assert r3["flag"][0] == 3


def test_smallest_t_frac_wins():
    """two bounds crossed in the same step: the earlier crossing decides flag and end time (checks.jl:37-41)"""
    m = O.make_model("LCO")
    a, _ = _run(m, 4.0, 0.0, eta_plating_min=0.05)
    b, _ = _run(m, 4.0, 0.0, c_e_min=300.0)
    both, _ = _run(m, 4.0, 0.0, eta_plating_min=0.05, c_e_min=300.0)
    first = a if a["t_end"][0] < b["t_end"][0] else b
    assert both["flag"][0] == first["flag"][0]
    assert both["t_end"][0] == first["t_end"][0]
