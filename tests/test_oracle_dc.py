"""The concentration-rate inputs dc_s_p_max / dc_s_p_min / dc_s_n_max / dc_s_n_min / dc_e_max / dc_e_min
(src/physics_equations/input_methods.jl:190-245) in the CPU oracle: a run_residual  val - Y'[ind]  where ind is the
arg-max / arg-min surface concentration (or electrolyte concentration) of the previous solution's last point.
Nothing in the reference executes them; what can be pinned is the closed form the control equation implies: the chosen
state moves at exactly the requested rate, Y[ind](t) = Y[ind](0) + val * t, and `:hold` keeps it where it was."""
import numpy as np
import pytest

import oracle as O

KINDS = ("dc_s_p_max", "dc_s_p_min", "dc_s_n_max", "dc_s_n_min", "dc_e_max", "dc_e_min")


def _target(m, L, Y, kind):
    if kind.startswith("dc_e"):
        v = Y[L.c_e:L.c_e + L.Nx]
        return L.c_e + int(v.argmax() if kind.endswith("max") else v.argmin())
    n = kind[5] == "n"
    base, cnt, Nr = (L.c_s_n, m.N_n, m.N_r_n) if n else (L.c_s_p, m.N_p, m.N_r_p)
    v = Y[base + Nr - 1:base + cnt * Nr:Nr]
    return base + Nr - 1 + Nr * int(v.argmax() if kind.endswith("max") else v.argmin())


@pytest.mark.parametrize("temperature", [False, True])
def test_the_chosen_state_moves_at_the_requested_rate(temperature):
    m = O.make_model("LCO", temperature=temperature)
    L = O.layout(m)
    th = O.theta_defaults()[None, :]
    tol = dict(reltol=1e-8, abstol=1e-8, reltol_init=1e-8, abstol_init=1e-8)
    r0 = O.simulate_batch(m, th, O.make_run("I", -1.0, tf=600.0), O.default_opts(**tol), O.default_bounds())
    st = r0["state"]
    Y, YP = st["Y"][0], st["YP"][0]
    for kind in KINDS:
        idx = _target(m, L, Y, kind)
        for val, ik in ((0.5 * YP[idx], "value"), (123.0, "hold")):
            r = O.simulate_batch(m, th, O.make_run(kind, val, tf=300.0, input_kind=ik, new_run=False), O.default_opts(**tol),
                                 O.default_bounds(), state=st)
            assert r["flag"][0] == 0
            want = 0.0 if ik == "hold" else val * 300.0          # `:hold` is rate 0 whatever number rides along
            got = r["state"]["Y"][0][idx] - Y[idx]
            assert abs(got - want) <= 1e-6 * max(abs(want), 1.0) + 1e-7 * abs(Y[idx]), (kind, ik, got, want)
            # the current is whatever it takes: it moved away from the -1 of the first segment (in |I|, for a halved rate)
            if ik == "value" and kind.startswith("dc_s"):
                assert abs(r["I_end"][0]) < 0.95


def test_needs_a_previous_solution_and_its_jacobian_row():
    m = O.make_model("LCO")
    L = O.layout(m)
    th = O.theta_defaults()
    r = O.simulate_batch(m, th[None, :], O.make_run("dc_e_max", 0.0, tf=10.0), O.default_opts(), O.default_bounds())
    assert r["flag"][0] < 0                                      # @assert !isempty(sol.Y), input_methods.jl:196
    # the control row: -gamma on the chosen column and nothing else; inside newtons_method! it is the j entry of the
    # chosen state's own differential row
    Y = O.initial_guess(m, th, 0.5); Y[L.I] = 1.0
    it, Y, YP = O.newton_init(m, th, O.make_run("I", 1.0), O.default_opts(), Y)
    for idx in (L.c_s_n + 9 + 30, L.c_e + 4, L.c_e + 14):
        run = O.make_run(("dc", idx), 0.0)
        J = O.jacobian(m, th, run, 0.0, Y, YP, 0.37)
        res = O.residual(m, th, run, 0.0, Y, YP)
        assert res[L.I] == -YP[idx]
        h = 1e-2 * max(abs(YP[idx]), 1.0)
        YPp = YP.copy(); YPp[idx] += h
        assert abs((O.residual(m, th, run, 0.0, Y, YPp)[L.I] - res[L.I]) / h + 1.0) < 1e-8
        # complex-step Jacobian of the oracle: the row sums to -gamma
        assert abs(J.sum() - O.jacobian(m, th, O.make_run("I", 0.0), 0.0, Y, YP, 0.37).sum() + 1.0 + 0.37) < 1e-6 * np.abs(J).max()
