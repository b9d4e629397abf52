#!/usr/bin/env python3
"""Build-time code generator (product code): sympy -> CUDA device functions.

The reference generates its residual/Jacobian symbolically with Symbolics.jl
(/root/reference/src/generate_functions.jl:102-164, 289-325).  The B200 build keeps that idea
but only for the *constitutive laws* (the node-local nonlinear functions and their exact
symbolic derivatives); the finite-volume stencil assembly around them is hand-written
warp-parallel CUDA (csrc/plb_device.cuh).  Emitted file: csrc/laws_generated.cuh

Laws restated from /root/reference/src/physics_equations/custom_functions.jl:
  K_eff (:96), D_eff (:83), OCV_LCO (:123-136), OCV_LiC6 (:139-152), OCV_NMC (:154-162),
  OCV_LiC6_with_NMC (:164-174)
and the Fickian finite-difference particle operator from
  /root/reference/src/physics_equations/numerical_tools.jl:8-87 and residuals.jl:128-180.
"""
import os
import sys

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "csrc", "laws_generated.cuh")


from sympy.printing.c import C99CodePrinter


class _Printer(C99CodePrinter):
    """integer powers as multiplications (no pow() calls in device code)"""

    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer and 2 <= abs(int(e)) <= 4:
            bs = self._print(b)
            if not (b.is_Symbol or b.is_Number):
                bs = "(" + bs + ")"
            prod = "*".join([bs] * abs(int(e)))
            return "(" + prod + ")" if e > 0 else "(1.0/(" + prod + "))"
        return super()._print_Pow(expr)

    def _print_Function(self, expr):
        if expr.func.__name__ == "__drcp_rn":
            return "__drcp_rn(" + self._print(expr.args[0]) + ")"
        return super()._print_Function(expr)


_printer = _Printer()


def ccode(e):
    return _printer.doprint(e)


def rat(x):
    """exact rational of a decimal literal as written in the reference source"""
    return sp.Rational(str(x))


_rcp = sp.Function("__drcp_rn")


def _div_to_rcp(e):
    """a/b -> a*__drcp_rn(b): a double-precision division is ~20 instructions, the correctly rounded
    reciprocal + a multiply about half of that (the result differs from the division by at most 1 ulp)"""
    return e.replace(lambda x: x.is_Pow and x.exp.is_Integer and x.exp.is_negative,
                     lambda x: _rcp(x.base ** (-x.exp)))


def emit_fn(name, args, outs, doc):
    """outs: list of (name, expr).  CSE + C code.  All args/outs double."""
    exprs = [e for _, e in outs]
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("t"), optimizations="basic")
    repl = [(s_, _div_to_rcp(e)) for s_, e in repl]
    red = [_div_to_rcp(e) for e in red]
    lines = [f"// {doc}", "__device__ __forceinline__ void " + name + "(" +
             ", ".join(f"const double {a}" for a in args) + ", " +
             ", ".join(f"double& {o}" for o, _ in outs) + ") {"]
    for s, e in repl:
        lines.append(f"    const double {s} = {ccode(e)};")
    for (o, _), e in zip(outs, red):
        lines.append(f"    {o} = {ccode(e)};")
    lines.append("}")
    return "\n".join(lines)


def hornerize(expr, var):
    num, den = sp.fraction(sp.together(expr))
    return sp.horner(sp.expand(num), var) / sp.horner(sp.expand(den), var)


def poly_eval(coeffs, x):
    return sum(rat(c) * x**i for i, c in enumerate(coeffs))


NR_BUILT = (10, 12, 14)


def emit_particle_operator(out, n):
    """MC (combined first/second-derivative stencil of a particle with n radial nodes), its structural mask and its
    eigen-decomposition, inside namespace nr<n>."""
    M1 = np.zeros((n, n)); M2 = np.zeros((n, n))
    fb1 = np.array([[-109584.0, 322560, -564480, 752640, -705600, 451584, -188160, 46080, -5040],
                    [-5040.0, -64224, 141120, -141120, 117600, -70560, 28224, -6720, 720],
                    [720.0, -11520, -38304, 80640, -50400, 26880, -10080, 2304, -240],
                    [-240.0, 2880, -20160, -18144, 50400, -20160, 6720, -1440, 144]])
    mid1 = np.array([144.0, -1536, 8064, -32256, 0, 32256, -8064, 1536, -144])
    lb1 = np.array([[-144.0, 1440, -6720, 20160, -50400, 18144, 20160, -2880, 240],
                    [240.0, -2304, 10080, -26880, 50400, -80640, 38304, 11520, -720],
                    [-720.0, 6720, -28224, 70560, -117600, 141120, -141120, 64224, 5040],
                    [5040.0, -46080, 188160, -451584, 705600, -752640, 564480, -322560, 109584]])
    M1[:4, :9] = fb1
    for k, i in enumerate(range(4, n - 4)):
        M1[i, k:k + 9] = mid1
    M1[n - 4:, n - 9:] = lb1
    fb2 = np.array([[-415 / 6, 96, -36, 32 / 3, -3 / 2, 0], [10.0, -15, -4, 14, -6, 1]])
    mid2 = np.array([-1.0, 16, -30, 16, -1])
    lb2 = np.array([[1.0, -6, 14, -4, -15, 10], [0.0, -3 / 2, 32 / 3, -36, 96, -415 / 6]])
    M2[:2, :6] = fb2
    for k, i in enumerate(range(2, n - 2)):
        M2[i, k:k + 5] = mid2
    M2[n - 2:, n - 6:] = lb2
    dr = 1.0 / (n - 1)
    c1 = 1.0 / (40320 * dr); c2 = 1.0 / (12 * dr * dr)
    Mc = np.zeros((n, n))
    Mc[0] = 3 * c2 * M2[0]
    for r in range(1, n - 1):
        rk = r / (n - 1)
        Mc[r] = c2 * M2[r] + (2.0 / rk) * c1 * M1[r]
    Mc[n - 1] = c2 * M2[n - 1]
    bj = 50 * dr * c2 + 2.0  # multiplies d1_bc = -j*Rp/D_s in the surface row
    # the structural pattern must not lose an entry to a floating-point cancellation of the two stencils
    struct = (M1 != 0.0) | (M2 != 0.0)
    struct[0] = M2[0] != 0.0; struct[n - 1] = M2[n - 1] != 0.0
    assert np.array_equal(Mc != 0.0, struct), n
    out.append(f"namespace nr{n} {{")
    out.append(f"// combined Fickian operator for N_r = {n}:  rhs_cs = kappa*(MC*c_s + BJ*d1bc e_surf),  kappa = D_s/Rp^2")
    out.append(f"constexpr int NR = {n};")
    out.append("static __device__ __constant__ double MC[NR][NR] = {")
    for r in range(n):
        out.append("    {" + ", ".join(repr(float(v)) for v in Mc[r]) + "},")
    out.append("};")
    out.append(f"constexpr double BJ = {bj!r};")
    # structural pattern of the particle block (which (r,c) are non-zero), as a bit mask per row
    masks = [sum(1 << cc for cc in range(n) if Mc[r, cc] != 0.0) for r in range(n)]
    out.append("__host__ __device__ constexpr unsigned mc_mask(int r) {\n    return " +
               " ".join(f"r == {r} ? {hex(v)}u :" for r, v in enumerate(masks)) + " 0u;\n}")
    # eigen-decomposition MC = EV diag(EL) EVI (real spectrum, cond(EV) ~ 60): with a node-dependent
    # D_s(T) the particle block kappa_x*MC - cj*I is different at every node, but it is diagonal in
    # this fixed basis, so no per-node factorisation is needed (thermal variant).
    import mpmath as mp
    mp.mp.dps = 60
    A = mp.matrix(Mc.tolist())
    E, ER = mp.eig(A)
    assert all(abs(mp.im(e)) < mp.mpf(10) ** -40 for e in E), "particle operator has complex eigenvalues"
    order = sorted(range(n), key=lambda i: float(mp.re(E[i])))
    EV = mp.matrix(n, n)
    for k, i in enumerate(order):
        col = [mp.re(ER[r, i]) for r in range(n)]
        nrm = mp.sqrt(sum(v * v for v in col))
        for r in range(n):
            EV[r, k] = col[r] / nrm
    EVI = EV ** -1
    EL = [mp.re(E[i]) for i in order]
    if abs(EL[-1]) < mp.mpf(10) ** -30:
        EL[-1] = mp.mpf(0)          # conservation: MC * 1 = 0 exactly
    chk = max(abs((EV * mp.diag(EL) * EVI - A)[r, cc]) for r in range(n) for cc in range(n))
    assert chk < mp.mpf(10) ** -40, chk

    def arr(name, Mx):
        out.append(f"static __device__ __constant__ double {name}[NR][NR] = {{")
        for r in range(n):
            out.append("    {" + ", ".join(repr(float(Mx[r, cc])) for cc in range(n)) + "},")
        out.append("};")
    out.append("// MC = EV * diag(EL) * EVI (mpmath, 60 digits, rounded to double)")
    arr("EV", EV)
    arr("EVI", EVI)
    out.append("static __device__ __constant__ double EL[NR] = {" + ", ".join(repr(float(v)) for v in EL) + "};")
    out.append(f"}}  // namespace nr{n}")
    print("N_r", n, "particle nnz", sum(bin(v).count("1") for v in masks))


def emit_spectral_operator(out, n):
    """Fickian_method = :spectral ("BETA", residuals.jl:181-235): Chebyshev collocation.  With dc = P D R c (rows of the two
    boundary conditions removed by P, R = the centre->surface reversal) and dc[surface] = -j Rp / (2 D_s), the reference's
    right-hand side is linear:  rhs_cs = kappa * (MC c + GJ * d1bc),  kappa = D_s/Rp^2, d1bc = -j Rp/D_s -- the same form
    as the finite-difference scheme, with a DENSE constant block MC = 4 L1 P D R and a FULL coupling vector GJ = 2 L1 e_1.
    Formed in 60-digit arithmetic, rounded to double; namespace sp<n>."""
    import mpmath as mp
    mp.mp.dps = 60
    N = n - 1
    x = [mp.cos(mp.pi * k / N) for k in range(n)]
    cc = [(2 if k in (0, N) else 1) * (-1) ** k for k in range(n)]
    D = mp.matrix(n, n)
    for i in range(n):
        for k in range(n):
            D[i, k] = (mp.mpf(cc[i]) / cc[k]) / ((x[i] - x[k]) + (1 if i == k else 0))
    for i in range(n):
        D[i, i] -= sum(D[i, k] for k in range(n))
    R = mp.matrix(n, n)
    for i in range(n):
        R[i, n - 1 - i] = 1
    P = mp.eye(n); P[0, 0] = 0; P[n - 1, n - 1] = 0
    W = mp.diag([(xi + 1) ** 2 for xi in x])
    Lnum = R * (D * W)
    L1 = mp.matrix(n, n)
    for q in range(n):
        L1[0, q] = 3 * D[n - 1, q]                     # L'Hopital at the centre
    for r in range(1, n):
        for q in range(n):
            L1[r, q] = Lnum[r, q] / (x[n - 1 - r] + 1) ** 2
    A = 4 * L1 * P * D * R
    g = [2 * L1[r, 0] for r in range(n)]
    E, ER = mp.eig(A)
    assert all(abs(mp.im(e)) < mp.mpf(10) ** -40 for e in E), "spectral particle operator has complex eigenvalues"
    order = sorted(range(n), key=lambda i: float(mp.re(E[i])))
    EV = mp.matrix(n, n)
    for k, i in enumerate(order):
        col = [mp.re(ER[r, i]) for r in range(n)]
        nrm = mp.sqrt(sum(v * v for v in col))
        for r in range(n):
            EV[r, k] = col[r] / nrm
    EVI = EV ** -1
    EL = [mp.re(E[i]) for i in order]
    if abs(EL[-1]) < mp.mpf(10) ** -30:
        EL[-1] = mp.mpf(0)          # conservation: MC * 1 = 0
    chk = max(abs((EV * mp.diag(EL) * EVI - A)[r, q]) for r in range(n) for q in range(n))
    assert chk < mp.mpf(10) ** -35, chk
    EVIG = [sum(EVI[i, q] * g[q] for q in range(n)) for i in range(n)]
    out.append(f"namespace sp{n} {{")
    out.append(f"// spectral (Chebyshev) Fickian operator for N_r = {n}:  rhs_cs = kappa*(MC*c_s + GJ*d1bc),  kappa = D_s/Rp^2")
    out.append(f"constexpr int NR = {n};")
    out.append("constexpr double BJ = 1.0;      // (the j coupling is GJ[r] on every row)")
    out.append("__host__ __device__ constexpr unsigned mc_mask(int) { return " + hex((1 << n) - 1) + "u; }   // dense block")

    def arr(name, Mx):
        out.append(f"static __device__ __constant__ double {name}[NR][NR] = {{")
        for r in range(n):
            out.append("    {" + ", ".join(repr(float(Mx[r, q])) for q in range(n)) + "},")
        out.append("};")

    def vec(name, v):
        out.append(f"static __device__ __constant__ double {name}[NR] = {{" + ", ".join(repr(float(t)) for t in v) + "};")
    arr("MC", A)
    vec("GJ", g)
    out.append("// MC = EV * diag(EL) * EVI ;  EVIG = EVI * GJ")
    arr("EV", EV)
    arr("EVI", EVI)
    vec("EL", EL)
    vec("EVIG", EVIG)
    out.append(f"}}  // namespace sp{n}")
    print("spectral N_r", n, "cond(EV)", float(mp.norm(EV, 2) * mp.norm(EVI, 2)))


def main():
    th, c, T = sp.symbols("th c T", real=True)
    out = []
    out.append("// GENERATED by petlion.jl_b200/codegen/gen_laws.py -- do not edit.\n#pragma once\n"
               "namespace plb { namespace laws {\n")

    # ---- K_eff(c_e, T), custom_functions.jl:96 -------------------------------------------
    a = ((rat("-10.5") + rat("0.668") * rat("1e-3") * c + rat("0.494") * rat("1e-6") * c**2)
         + (rat("0.074") - rat("1.78") * rat("1e-5") * c - rat("8.86") * rat("1e-10") * c**2) * T
         + (rat("-6.96") * rat("1e-5") + rat("2.8") * rat("1e-8") * c) * T**2)
    # the reference evaluates the products of literals in Float64; keep them as doubles
    a = sp.N(a, 17)
    a = sp.collect(sp.expand(a), c)
    K = sp.Float("1e-4") * c * a**2
    a_h = sp.horner(sp.expand(a), c)
    A = sp.Symbol("A")
    out.append(emit_fn("K_eff", ["c", "T"],
                       [("K", sp.Float("1e-4") * c * a_h**2),
                        ("dKdc", sp.Float("1e-4") * a_h**2 + sp.Float("2e-4") * c * a_h * sp.horner(sp.expand(sp.diff(a, c)), c))],
                       "K_eff(c_e,T) and dK/dc_e  (custom_functions.jl:96)"))
    # thermal variant: also dK/dT (the T state is an unknown when temperature=true)
    a_T = sp.horner(sp.expand(sp.diff(a, T)), c)
    out.append(emit_fn("K_eff_T", ["c", "T"],
                       [("K", sp.Float("1e-4") * c * a_h**2),
                        ("dKdc", sp.Float("1e-4") * a_h**2 + sp.Float("2e-4") * c * a_h * sp.horner(sp.expand(sp.diff(a, c)), c)),
                        ("dKdT", sp.Float("2e-4") * c * a_h * a_T)],
                       "K_eff(c_e,T), dK/dc_e and dK/dT  (custom_functions.jl:96)"))

    # ---- D_eff(c_e, T), custom_functions.jl:83 --------------------------------------------
    e = (sp.Float("-4.43") - sp.Float("54.0") / (T - 229 - sp.Float("5e-3") * c) - sp.Float("0.22e-3") * c)
    D = sp.Float("1e-4") * sp.exp(e * sp.log(10).evalf(17))
    out.append(emit_fn("D_eff_nl", ["c", "T"], [("D", D), ("dDdc", sp.diff(D, c))],
                       "D_eff(c_e,T) and dD/dc_e  (custom_functions.jl:83)"))

    # ---- OCV_LCO, custom_functions.jl:123-136 ---------------------------------------------
    x = sp.Symbol("x")  # x = th^2
    num = (sp.Float("-4.656") + sp.Float("88.669") * x - sp.Float("401.119") * x**2 + sp.Float("342.909") * x**3
           - sp.Float("462.471") * x**4 + sp.Float("433.434") * x**5)
    den = (-1 + sp.Float("18.933") * x - sp.Float("79.532") * x**2 + sp.Float("37.311") * x**3
           - sp.Float("73.083") * x**4 + sp.Float("95.96") * x**5)
    nh, dh = sp.horner(num, x), sp.horner(den, x)
    dnh, ddh = sp.horner(sp.diff(num, x), x), sp.horner(sp.diff(den, x), x)
    U = (nh / dh).subs(x, th**2)
    dU = ((dnh * dh - nh * ddh) / dh**2).subs(x, th**2) * 2 * th
    nT = (sp.Float("0.199521039") - sp.Float("0.928373822") * th + sp.Float("1.364550689000003") * th**2
          - sp.Float("0.6115448939999998") * th**3)
    dT = (1 - sp.Float("5.661479886999997") * th + sp.Float("11.47636191") * th**2
          - sp.Float("9.82431213599998") * th**3 + sp.Float("3.048755063") * th**4)
    nTh, dTh = sp.horner(nT, th), sp.horner(dT, th)
    dUdT = sp.Float("-0.001") * nTh / dTh
    ddUdT = sp.Float("-0.001") * (sp.horner(sp.diff(nT, th), th) * dTh - nTh * sp.horner(sp.diff(dT, th), th)) / dTh**2
    out.append(emit_fn("OCV_LCO", ["th"], [("U", U), ("dU", dU), ("dUdT", dUdT), ("ddUdT", ddUdT)],
                       "OCV_LCO: U(th), dU/dth, dU/dT(th), d(dU/dT)/dth  (custom_functions.jl:123-136)"))
    out.append(emit_fn("OCV_LCO_U", ["th"], [("U", U), ("dU", dU)],
                       "OCV_LCO without the entropic term (isothermal run at T == T_ref: temperature_switch, custom_functions.jl:1)"))

    # ---- OCV_LiC6, custom_functions.jl:139-152 (physical branch th > 1e-4) ----------------
    s = sp.Symbol("s", positive=True)  # s = sqrt(th)
    U = (sp.Float("0.7222") + sp.Float("0.1387") * th + sp.Float("0.029") * s - sp.Float("0.0172") / th
         + sp.Float("0.0019") / (s * th) + sp.Float("0.2808") * sp.exp(sp.Float("0.9") - 15 * th)
         - sp.Float("0.7984") * sp.exp(sp.Float("0.4465") * th - sp.Float("0.4108")))
    dU = (sp.Float("0.1387") + sp.Float("0.029") / (2 * s) + sp.Float("0.0172") / th**2
          - sp.Rational(3, 2) * sp.Float("0.0019") / (s * th**2)
          - 15 * sp.Float("0.2808") * sp.exp(sp.Float("0.9") - 15 * th)
          - sp.Float("0.7984") * sp.Float("0.4465") * sp.exp(sp.Float("0.4465") * th - sp.Float("0.4108")))
    cn = [0.005269056, 3.299265709, -91.79325798, 1004.911008, -5812.278127, 19329.7549, -37147.8947,
          38379.18127, -16515.05308]
    cd = [1, -48.09287227, 1017.234804, -10481.80419, 59431.3, -195881.6488, 374577.3152, -385821.1607,
          165705.8597]
    nT = sum(sp.Float(repr(v)) * th**i for i, v in enumerate(cn))
    dT = sum(sp.Float(repr(v)) * th**i for i, v in enumerate(cd))
    nTh, dTh = sp.horner(nT, th), sp.horner(dT, th)
    dUdT = sp.Float("0.001") * nTh / dTh
    ddUdT = sp.Float("0.001") * (sp.horner(sp.diff(nT, th), th) * dTh - nTh * sp.horner(sp.diff(dT, th), th)) / dTh**2
    out.append(emit_fn("OCV_LiC6", ["th", "s"], [("U", U), ("dU", dU), ("dUdT", dUdT), ("ddUdT", ddUdT)],
                       "OCV_LiC6 (s = sqrt(th), th > 1e-4 branch of sqrt_ReLU)  (custom_functions.jl:139-152)"))
    out.append(emit_fn("OCV_LiC6_U", ["th", "s"], [("U", U), ("dU", dU)],
                       "OCV_LiC6 without the entropic term (isothermal run at T == T_ref)"))

    # ---- OCV_NMC / OCV_LiC6_with_NMC, custom_functions.jl:154-174 -------------------------
    U = sp.horner(sp.Float("-10.72") * th**4 + sp.Float("23.88") * th**3 - sp.Float("16.77") * th**2
                  + sp.Float("2.595") * th + sp.Float("4.563"), th)
    out.append(emit_fn("OCV_NMC", ["th"], [("U", U), ("dU", sp.horner(sp.diff(sp.expand(U), th), th))],
                       "OCV_NMC  (custom_functions.jl:154-162)"))
    U = (sp.Float("0.1493") + sp.Float("0.8493") * sp.exp(sp.Float("-61.79") * th)
         + sp.Float("0.3824") * sp.exp(sp.Float("-665.8") * th) - sp.exp(sp.Float("39.42") * th - sp.Float("41.92"))
         - sp.Float("0.03131") * sp.atan(sp.Float("25.59") * th - sp.Float("4.099"))
         - sp.Float("0.009434") * sp.atan(sp.Float("32.49") * th - sp.Float("15.74")))
    out.append(emit_fn("OCV_LiC6_NMC", ["th"], [("U", U), ("dU", sp.diff(U, th))],
                       "OCV_LiC6_with_NMC  (custom_functions.jl:164-174)"))

    # ---- NMC_LGM50 / LiC6_LGM50 (Chen et al. 2020), params.jl:557-573, 632-640, 648, 662 -------
    U = (sp.Float("-0.8090") * th + sp.Float("4.4875") - sp.Float("0.0428") * sp.tanh(sp.Float("18.5138") * (th - sp.Float("0.5542")))
         - sp.Float("17.7326") * sp.tanh(sp.Float("15.7890") * (th - sp.Float("0.3117")))
         + sp.Float("17.5842") * sp.tanh(sp.Float("15.9308") * (th - sp.Float("0.3120"))))
    out.append(emit_fn("OCV_NMC811", ["th"], [("U", U), ("dU", sp.diff(U, th))], "OCV of NMC_LGM50  (params.jl:557-562)"))
    U = (sp.Float("1.9793") * sp.exp(sp.Float("-39.3631") * th) + sp.Float("0.15561")
         - sp.Float("0.0909") * sp.tanh(sp.Float("29.8538") * (th - sp.Float("0.1234")))
         - sp.Float("0.04478") * sp.tanh(sp.Float("14.9159") * (th - sp.Float("0.2769")))
         - sp.Float("0.0205") * sp.tanh(sp.Float("30.4444") * (th - sp.Float("0.6103")))
         - sp.Float("0.09259") * sp.tanh(sp.Float("17.08") * (th - 1)))
    out.append(emit_fn("OCV_LiC6_LGM50", ["th"], [("U", U), ("dU", sp.diff(U, th))], "OCV of LiC6_LGM50  (params.jl:632-636)"))
    q = sp.Symbol("q", positive=True)      # q = sqrt(c/1000)
    ck = c / 1000
    K = sp.Float("0.1297") * ck**3 - sp.Float("2.51") * ck * q + sp.Float("3.329") * ck
    dK = (3 * sp.Float("0.1297") * ck**2 - sp.Float("1.5") * sp.Float("2.51") * q + sp.Float("3.329")) / 1000
    out.append(emit_fn("K_eff_LGM50", ["c", "q"], [("K", K), ("dKdc", dK)],
                       "K_eff_LGM50(c_e) and dK/dc_e, q = sqrt(c_e/1000)  (params.jl:662)"))
    D = ck**2 - sp.Float("4.516715942688196") * ck + sp.Float("5.5287696156470325")
    out.append(emit_fn("D_eff_LGM50", ["c"], [("D", D), ("dDdc", sp.diff(D, c))],
                       "D_eff_LGM50(c_e) / D_e and its c_e derivative  (params.jl:648)"))

    # ---- Fickian particle operator, numerical_tools.jl:8-87, residuals.jl:128-180: one table set per built N_r
    # (N_r = 10 is every parameter set's default, params.jl:134-136; the others are sibling builds, -DPLB_NR=n)
    for n in NR_BUILT:
        emit_particle_operator(out, n)
    emit_spectral_operator(out, 10)
    out.append("#ifndef PLB_NR\n#define PLB_NR 10\n#endif")
    out.append("#ifndef PLB_SPECTRAL\n#define PLB_SPECTRAL 0      // 1: Fickian_method = :spectral (sibling builds, N_r = 10)\n#endif")
    out.append("#define PLB_NR_NS2_(n) nr##n\n#define PLB_NR_NS_(n) PLB_NR_NS2_(n)")
    out.append("#if PLB_SPECTRAL\n#if PLB_NR != 10\n#error \"the spectral particle operator is generated for N_r = 10\"\n#endif\nusing namespace sp10;\n#else")
    out.append("using namespace PLB_NR_NS_(PLB_NR);   // laws::NR, laws::MC, laws::BJ, laws::mc_mask, laws::EV, laws::EVI, laws::EL\n#endif")
    out.append("\n}}  // namespace plb::laws\n")
    with open(OUT, "w") as f:
        f.write("\n".join(out))
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
