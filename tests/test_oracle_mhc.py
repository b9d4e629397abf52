"""rxn_MHC in the CPU oracle (custom_functions.jl:233-298).  Nothing in the reference executes this rate law, so what
can be pinned is the function itself: its closed forms at special arguments, and the complex-step Jacobian (whose erf
is a first-order expansion in the imaginary part) against central differences of the residual."""
import math

import numpy as np

import oracle as O

RX = dict(rxn_p="MHC", rxn_n="MHC")
F, R = 96485.3321233, 8.31446261815324


def _theta(lam_p=None, lam_n=None):
    th = O.theta_defaults()
    n = O.theta_names()
    if lam_p is not None:
        th[n.index("lambda_MHC_p")] = lam_p
        th[n.index("lambda_MHC_n")] = lam_n if lam_n is not None else lam_p
    return th, n


def _mhc(cs, ce, T, eta, k, lam, cmax, ce0):
    """the reference function, transcribed with math.erf"""
    etah = eta * F / (R * T)
    eta_f = etah + math.log(max(1e-4, (ce / ce0) / (cs / cmax)))
    a = 1.0 + math.sqrt(lam)
    k0 = k / ((1.0 - math.erf((lam - math.sqrt(a)) / (2.0 * math.sqrt(lam)))) / 2.0)
    coeff = k0 * (1.0 - math.erf((lam - math.sqrt(a + eta_f ** 2)) / (2.0 * math.sqrt(lam))))
    return coeff * (ce0 * cs / (1.0 + math.exp(-eta_f)) - ce * cmax / (1.0 + math.exp(eta_f))) * math.sqrt(max(0.0, (1.0 - cs / cmax) / ce0))


def _j_rows(m, th, Y, YP=None):
    L = O.layout(m)
    YP = np.zeros_like(Y) if YP is None else YP
    r = O.residual(m, th, O.make_run("I", 1.0), 0.0, Y, YP)
    Ne = m.N_p + m.N_n
    return r[L.j:L.j + Ne] + Y[L.j:L.j + Ne]      # res_j = j_calc - j


def test_rate_law_values():
    for lam in (6.26e-20, 15.0, 3.0):
        th, n = _theta(lam, 0.7 * lam if lam > 1 else lam)
        m = O.make_model("LCO", **RX)
        L = O.layout(m)
        Y = O.initial_guess(m, th, 0.6)
        rng = np.random.default_rng(3)
        Y[L.phi_s:L.phi_s + m.N_p] += 0.03 * rng.uniform(-1, 1, m.N_p)
        Y[L.c_e:L.c_e + L.Nx] *= 1 + 0.2 * rng.uniform(-1, 1, L.Nx)
        jc = _j_rows(m, th, Y)
        g = dict(zip(n, th))
        for e in range(m.N_p + m.N_n):
            isp = e < m.N_p
            x = e if isp else e + m.N_s
            cs = Y[(L.c_s_p if isp else L.c_s_n) + (e if isp else e - m.N_p) * 10 + 9]
            cmax = g["c_max_p" if isp else "c_max_n"]
            # OCV through the oracle's own BV model: eta = (asinh of the BV rate) -- instead, rebuild eta from a BV twin
            mb = O.make_model("LCO")
            jb = _j_rows(mb, th, Y)[e]
            k = g["k_p" if isp else "k_n"]
            arg = Y[L.c_e + x] * cs * (cmax - cs)
            eta = math.asinh(jb / (2 * k * math.sqrt(arg))) * 2 * R * g["T0"] / F
            lam_e = g["lambda_MHC_p" if isp else "lambda_MHC_n"]
            ref = _mhc(cs, Y[L.c_e + x], g["T0"], eta, k, lam_e, cmax, g["c_e0"])
            assert abs(jc[e] - ref) <= 1e-9 * abs(ref) + 1e-18, (lam, e, jc[e], ref)
            if lam < 1e-10:
                # the reference's default: both erf are -1 and the law is a logistic form with coefficient 2k
                etaf = eta * F / (R * g["T0"]) + math.log((Y[L.c_e + x] / g["c_e0"]) / (cs / cmax))
                closed = 2 * k * (g["c_e0"] * cs / (1 + math.exp(-etaf)) - Y[L.c_e + x] * cmax / (1 + math.exp(etaf))) * math.sqrt((1 - cs / cmax) / g["c_e0"])
                assert abs(jc[e] - closed) <= 1e-12 * abs(closed) + 1e-20


def test_jacobian_against_central_differences():
    for temperature in (False, True):
        th, n = _theta(15.0, 9.0)
        m = O.make_model("LCO", temperature=temperature, **RX)
        L = O.layout(m)
        run = O.make_run("I", 1.0)
        Y = O.initial_guess(m, th, 0.45)
        Y[L.I] = 1.0
        it, Y, YP = O.newton_init(m, th, run, O.default_opts(), Y)
        assert it > 0
        gam = 0.37
        J = O.jacobian(m, th, run, 0.0, Y, YP, gam)
        cp, rv = O.jac_pattern(m, "I")
        Ne = m.N_p + m.N_n
        rows = set(range(L.j, L.j + Ne))
        worst = 0.0
        for c in range(L.N_tot):
            ks = [k for k in range(cp[c], cp[c + 1]) if rv[k] in rows]
            if not ks:
                continue
            h = 1e-6 * max(abs(Y[c]), 1e-3)
            Yp_, Ym_ = Y.copy(), Y.copy()
            Yp_[c] += h; Ym_[c] -= h
            # J = dF/dY + gamma dF/dY': the j rows are algebraic, no Y' part
            d = (O.residual(m, th, run, 0.0, Yp_, YP) - O.residual(m, th, run, 0.0, Ym_, YP)) / (2 * h)
            for k in ks:
                scale = max(abs(J[k]), abs(d[rv[k]]), 1e-30)
                worst = max(worst, abs(J[k] - d[rv[k]]) / scale)
        assert worst < 1e-6, worst


def test_discharge_runs_and_differs_from_butler_volmer():
    th, n = _theta(15.0, 12.0)
    out = {}
    for name, rx in (("bv", {}), ("mhc", RX)):
        m = O.make_model("LCO", **rx)
        r = O.simulate_batch(m, th[None, :], O.make_run("I", -1.0, tf=1e6), O.default_opts(), O.default_bounds())
        out[name] = r
        assert r["flag"][0] == 3 and abs(r["t_end"][0] - 3600.0) < 1e-6        # SOC_min at 1C: the capacity is the same
    assert 1e-3 < abs(out["bv"]["V_end"][0] - out["mhc"]["V_end"][0]) < 0.1
