"""Lane-level numpy prototype of the structured Newton-matrix solver of the thermal variant
(petlion.jl_b200/csrc/plb_device.cuh, PLB_TH=1).  TEST INFRASTRUCTURE: it exists so that the algebra of
the CUDA code (node-local elimination with the eigen-basis particle solve, current-collector chains,
twisted 4x4 block-Thomas with the four two-node-wide temperature-row couplings absorbed exactly, and
the applied-current border) can be checked on the CPU against a dense solve of the oracle's Jacobian.

Every array has a leading "lane" axis of 32; `shfl(v, src)` is the warp shuffle.
"""
import re
import os

import numpy as np

NR = 10
HERE = os.path.dirname(os.path.abspath(__file__))


def _laws():
    src = open(os.path.join(HERE, "..", "petlion.jl_b200", "csrc", "laws_generated.cuh")).read()

    def arr(name, two_d=True):
        i = src.index(f"double {name}[")
        blk = src[src.index("{", i):src.index("};", i) + 1]
        if two_d:
            return np.array([[float(v) for v in row.split(",")] for row in re.findall(r"\{([^{}]*)\}", blk)])
        return np.array([float(v) for v in blk.strip("{}; \n").split(",")])
    return arr("MC"), arr("EV"), arr("EVI"), arr("EL", False)


MC, EV, EVI, EL = _laws()


def shfl(v, src):
    return v[np.asarray(src)]


def inv4(a):
    """adjugate inverse of [...,4,4] (same formula as the CUDA inv4x4)"""
    a = np.asarray(a)
    A = lambda r, c: a[..., r, c]
    s0 = A(0, 0) * A(1, 1) - A(1, 0) * A(0, 1); s1 = A(0, 0) * A(1, 2) - A(1, 0) * A(0, 2)
    s2 = A(0, 0) * A(1, 3) - A(1, 0) * A(0, 3); s3 = A(0, 1) * A(1, 2) - A(1, 1) * A(0, 2)
    s4 = A(0, 1) * A(1, 3) - A(1, 1) * A(0, 3); s5 = A(0, 2) * A(1, 3) - A(1, 2) * A(0, 3)
    c5 = A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3); c4 = A(2, 1) * A(3, 3) - A(3, 1) * A(2, 3)
    c3 = A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2); c2 = A(2, 0) * A(3, 3) - A(3, 0) * A(2, 3)
    c1 = A(2, 0) * A(3, 2) - A(3, 0) * A(2, 2); c0 = A(2, 0) * A(3, 1) - A(3, 0) * A(2, 1)
    det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0
    i = 1.0 / det
    b = np.empty_like(a)
    b[..., 0, 0] = (A(1, 1) * c5 - A(1, 2) * c4 + A(1, 3) * c3) * i
    b[..., 0, 1] = (-A(0, 1) * c5 + A(0, 2) * c4 - A(0, 3) * c3) * i
    b[..., 0, 2] = (A(3, 1) * s5 - A(3, 2) * s4 + A(3, 3) * s3) * i
    b[..., 0, 3] = (-A(2, 1) * s5 + A(2, 2) * s4 - A(2, 3) * s3) * i
    b[..., 1, 0] = (-A(1, 0) * c5 + A(1, 2) * c2 - A(1, 3) * c1) * i
    b[..., 1, 1] = (A(0, 0) * c5 - A(0, 2) * c2 + A(0, 3) * c1) * i
    b[..., 1, 2] = (-A(3, 0) * s5 + A(3, 2) * s2 - A(3, 3) * s1) * i
    b[..., 1, 3] = (A(2, 0) * s5 - A(2, 2) * s2 + A(2, 3) * s1) * i
    b[..., 2, 0] = (A(1, 0) * c4 - A(1, 1) * c2 + A(1, 3) * c0) * i
    b[..., 2, 1] = (-A(0, 0) * c4 + A(0, 1) * c2 - A(0, 3) * c0) * i
    b[..., 2, 2] = (A(3, 0) * s4 - A(3, 1) * s2 + A(3, 3) * s0) * i
    b[..., 2, 3] = (-A(2, 0) * s4 + A(2, 1) * s2 - A(2, 3) * s0) * i
    b[..., 3, 0] = (-A(1, 0) * c3 + A(1, 1) * c1 - A(1, 2) * c0) * i
    b[..., 3, 1] = (A(0, 0) * c3 - A(0, 1) * c1 + A(0, 2) * c0) * i
    b[..., 3, 2] = (-A(3, 0) * s3 + A(3, 1) * s1 - A(3, 2) * s0) * i
    b[..., 3, 3] = (A(2, 0) * s3 - A(2, 1) * s1 + A(2, 2) * s0) * i
    return b


class Geo:
    def __init__(self, Np=10, Ns=10, Nn=10, Na=10, Nz=10):
        self.Np, self.Ns, self.Nn, self.Na, self.Nz = Np, Ns, Nn, Na, Nz
        self.Nx, self.Ne = Np + Ns + Nn, Np + Nn
        self.off_cs = self.Nx
        self.off_T = self.off_cs + NR * self.Ne
        self.off_j = self.off_T + Na + self.Nx + Nz
        self.off_pe = self.off_j + self.Ne
        self.off_ps = self.off_pe + self.Nx
        self.off_I = self.off_ps + self.Ne
        self.N = self.off_I + 1
        self.mid = Np + Ns // 2
        lane = np.arange(32)
        self.lane = lane
        self.act = lane < self.Nx
        self.sec = np.where(lane < Np, 0, np.where(lane < Np + Ns, 1, np.where(lane < self.Nx, 2, 3)))
        self.elec = (self.sec == 0) | (self.sec == 2)
        self.e = np.where(self.sec == 0, lane, np.where(self.sec == 2, lane - Ns, -1))
        self.cha = lane < Na
        self.chz = (lane >= self.Nx - Nz) & (lane < self.Nx)
        # index of this lane's unknowns in the reference layout (-1: none)
        self.i_ce = np.where(self.act, lane, -1)
        self.i_pe = np.where(self.act, self.off_pe + lane, -1)
        self.i_T = np.where(self.act, self.off_T + Na + lane, -1)
        self.i_j = np.where(self.elec, self.off_j + self.e, -1)
        self.i_ps = np.where(self.elec, self.off_ps + self.e, -1)
        self.i_cs = np.where(self.elec[:, None], self.off_cs + self.e[:, None] * NR + np.arange(NR)[None, :], -1)
        self.i_Tx = np.where(self.cha, self.off_T + lane, np.where(self.chz, self.off_T + Na + self.Nx + (lane - (self.Nx - Nz)), -1))


def lane_jac_from_dense(g, J, cj):
    """Pull the per-lane coefficient set the CUDA lane_eval produces (LaneJac) out of a dense Jacobian
    dF/dY + cj dF/dY' of the oracle.  Differential diagonals are returned WITHOUT the -cj term."""
    def at(r, c):
        ok = (r >= 0) & (c >= 0)
        return np.where(ok, J[np.where(ok, r, 0), np.where(ok, c, 0)], 0.0)
    L = g.lane
    nb = lambda idx, d: np.where((L + d >= 0) & (L + d < 32), idx[np.clip(L + d, 0, 31)], -1)
    o = {}
    o["ceL"], o["ceD"], o["ceU"] = at(g.i_ce, nb(g.i_ce, -1)), at(g.i_ce, g.i_ce) + cj * g.act, at(g.i_ce, nb(g.i_ce, 1))
    o["ce_j"] = at(g.i_ce, g.i_j)
    surf = g.i_cs[:, NR - 1]
    o["j_cs"], o["j_ce"], o["j_pe"], o["j_ps"], o["j_T"] = at(g.i_j, surf), at(g.i_j, g.i_ce), at(g.i_j, g.i_pe), at(g.i_j, g.i_ps), at(g.i_j, g.i_T)
    o["peL"], o["peD"], o["peU"] = at(g.i_pe, nb(g.i_pe, -1)), at(g.i_pe, g.i_pe), at(g.i_pe, nb(g.i_pe, 1))
    o["pcL"], o["pcD"], o["pcU"] = at(g.i_pe, nb(g.i_ce, -1)), at(g.i_pe, g.i_ce), at(g.i_pe, nb(g.i_ce, 1))
    o["peTL"], o["peTD"], o["peTU"] = at(g.i_pe, nb(g.i_T, -1)), at(g.i_pe, g.i_T), at(g.i_pe, nb(g.i_T, 1))
    o["pe_j"] = at(g.i_pe, g.i_j)
    # Phi_s neighbours: only inside the same electrode (nb() of i_ps is -1 in the separator)
    o["psL"], o["psD"], o["psU"] = at(g.i_ps, nb(g.i_ps, -1)), np.where(g.elec, at(g.i_ps, g.i_ps), 1.0), at(g.i_ps, nb(g.i_ps, 1))
    o["ps_j"], o["ps_I"] = at(g.i_ps, g.i_j), at(g.i_ps, np.full(32, g.off_I))
    o["cs_j"] = at(surf, g.i_j)
    o["csT"] = np.stack([at(g.i_cs[:, r], g.i_T) for r in range(NR)], axis=1)
    o["kap"] = np.where(g.elec, at(g.i_cs[:, 0], g.i_cs[:, 1]) / MC[0, 1], 0.0)
    # T row
    iTl = nb(g.i_T, -1).copy(); iTr = nb(g.i_T, 1).copy()
    o["T_TL"] = at(g.i_T, iTl); o["T_TU"] = at(g.i_T, iTr)
    # chain couplings of the two end nodes
    o["T_chain"] = np.zeros(32)
    o["T_chain"][0] = J[g.i_T[0], g.off_T + g.Na - 1]
    o["T_chain"][g.Nx - 1] = J[g.i_T[g.Nx - 1], g.off_T + g.Na + g.Nx]
    o["T_TD"] = at(g.i_T, g.i_T) + cj * g.act
    o["T_j"], o["T_cs"] = at(g.i_T, g.i_j), at(g.i_T, surf)
    for nm, idx in (("T_ce", g.i_ce), ("T_pe", g.i_pe), ("T_ps", g.i_ps)):
        o[nm] = np.stack([at(g.i_T, nb(idx, d)) for d in (-2, -1, 0, 1, 2)], axis=1)
    # chains
    ix = g.i_Tx
    left = np.where(g.cha, nb(ix, -1), np.where(g.chz, np.where(L == g.Nx - g.Nz, g.i_T[g.Nx - 1], nb(ix, -1)), -1))
    right = np.where(g.cha, np.where(L == g.Na - 1, g.i_T[0], nb(ix, 1)), np.where(g.chz, nb(ix, 1), -1))
    left = np.where(g.cha & (L == 0), -1, left)
    right = np.where(g.chz & (L == g.Nx - 1), -1, right)
    o["Tx_L"], o["Tx_U"] = at(ix, left), at(ix, right)
    o["Tx_D"] = at(ix, ix) + cj * (ix >= 0)
    o["Tx_I"] = at(ix, np.full(32, g.off_I))
    o["g_ps0"], o["g_psN"], o["g_I"] = J[g.off_I, g.off_ps], J[g.off_I, g.off_ps + g.Ne - 1], J[g.off_I, g.off_I]
    return o


class Factor:
    pass


def core_solve(g, Fa, rb, rx):
    """(block rhs [32,4], chain rhs [32]) -> (u [32,4], ux [32]) for the matrix without the border."""
    L = g.lane
    nch = max(g.Na, g.Nz)
    # 1. chains, forward
    yx = rx.copy()
    for _ in range(nch - 1):
        yx = rx - Fa.chm * shfl(yx, Fa.cpred)
    # 2. fold into the two end nodes
    rb = rb.copy()
    tail = np.where(L == 0, g.Na - 1, g.Nx - g.Nz)
    rb[:, 3] -= Fa.hm * shfl(yx, tail)
    # 3. twisted block-Thomas
    r = rb.copy()
    y = r.copy()
    for it in range(Fa.n_in):
        if it == Fa.posA - 1 or it == Fa.posB - 1:
            mine = (Fa.pos == it + 1) & Fa.hasF
            y2 = shfl(y, Fa.pred2)
            r[:, 3] -= np.where(mine, np.einsum("lc,lc->l", Fa.Fr, y2), 0.0)
        a = shfl(y, Fa.pred)
        y = r - np.einsum("lrc,lc->lr", Fa.Wm, a)
    # (a special lane that is finalised by the very last inward iteration is not allowed: asserted in factor)
    a = shfl(y, np.full(32, g.mid - 1)); b = shfl(y, np.full(32, g.mid + 1))
    ym = r - np.einsum("lrc,lc->lr", Fa.Wm, a) - np.einsum("lrc,lc->lr", Fa.Pm, b)
    y = np.where(Fa.is_mid[:, None], ym, y)
    c = np.einsum("lrc,lc->lr", Fa.Dinv, y)
    P = np.where(Fa.is_mid[:, None, None], 0.0, Fa.Pm)
    u = c.copy()
    for it in range(Fa.n_out):
        a = shfl(u, Fa.succ)
        u = c - np.einsum("lrc,lc->lr", P, a)
    # chain heads: rows of node 0 / Nx-1 also reach two nodes ahead (T row only)
    a2 = shfl(u, Fa.succ2)
    t = np.einsum("lc,lc->l", Fa.Eo, a2[:, :3])
    u = u - Fa.Dinv[:, :, 3] * t[:, None]
    # 4. chains, backward
    head = np.where(g.cha, 0, g.Nx - 1)
    uT = shfl(u[:, 3], head)
    ux = yx * Fa.chip
    for _ in range(nch):
        us = np.where(Fa.is_tail, uT, shfl(ux, Fa.csucc))
        ux = (yx - Fa.chup * us) * Fa.chip
    ux = np.where(g.cha | g.chz, ux, 0.0)
    return u, ux


def factor(g, Jc, cj):
    L = g.lane
    Nx, mid = g.Nx, g.mid
    Fa = Factor()
    el = g.elec
    # ---- 1. particles in the eigen-basis of MC ---------------------------------------------------------
    pd = np.where(el[:, None], 1.0 / (Jc["kap"][:, None] * EL[None, :] - cj), 0.0)
    beta = Jc["cs_j"] * np.einsum("i,li,i->l", EV[NR - 1], pd, EVI[:, NR - 1])
    wT = np.einsum("ic,lc->li", EVI, Jc["csT"])
    tau = np.einsum("i,li,li->l", EV[NR - 1], pd, wT)
    Fa.pd, Fa.wT, Fa.csj = pd, wT, Jc["cs_j"]
    # ---- 2. node-local elimination of c_s and j --------------------------------------------------------
    inv_den = np.where(el, 1.0 / (-1.0 - Jc["j_cs"] * beta), 0.0)
    q = np.stack([-Jc["j_ce"] * inv_den, -Jc["j_pe"] * inv_den, -Jc["j_ps"] * inv_den,
                  -(Jc["j_T"] - Jc["j_cs"] * tau) * inv_den], axis=1)
    sj = np.stack([Jc["ce_j"], Jc["pe_j"], Jc["ps_j"], Jc["T_j"] - Jc["T_cs"] * beta], axis=1)
    Fa.q, Fa.inv_den, Fa.jcs, Fa.sj, Fa.tcs = q, inv_den, Jc["j_cs"], sj, Jc["T_cs"]
    D = np.zeros((32, 4, 4))
    D[:, 0, 0] = Jc["ceD"] - cj
    D[:, 1, 0], D[:, 1, 1], D[:, 1, 3] = Jc["pcD"], Jc["peD"], Jc["peTD"]
    D[:, 2, 2] = Jc["psD"]
    D[:, 3, 0], D[:, 3, 1], D[:, 3, 2] = Jc["T_ce"][:, 2], Jc["T_pe"][:, 2], Jc["T_ps"][:, 2]
    D[:, 3, 3] = Jc["T_TD"] - cj - Jc["T_cs"] * tau
    D += sj[:, :, None] * q[:, None, :]
    D[~g.act] = np.eye(4)

    def nine(c00, c10, c11, c13, c22, c30, c31, c32, c33):
        M = np.zeros((32, 4, 4))
        M[:, 0, 0], M[:, 1, 0], M[:, 1, 1], M[:, 1, 3], M[:, 2, 2] = c00, c10, c11, c13, c22
        M[:, 3, 0], M[:, 3, 1], M[:, 3, 2], M[:, 3, 3] = c30, c31, c32, c33
        return M
    L9 = nine(Jc["ceL"], Jc["pcL"], Jc["peL"], Jc["peTL"], Jc["psL"], Jc["T_ce"][:, 1], Jc["T_pe"][:, 1], Jc["T_ps"][:, 1], Jc["T_TL"])
    U9 = nine(Jc["ceU"], Jc["pcU"], Jc["peU"], Jc["peTU"], Jc["psU"], Jc["T_ce"][:, 3], Jc["T_pe"][:, 3], Jc["T_ps"][:, 3], Jc["T_TU"])
    L9[(~g.act) | (L == 0)] = 0.0
    U9[(~g.act) | (L >= Nx - 1)] = 0.0
    E2m = np.stack([Jc["T_ce"][:, 0], Jc["T_pe"][:, 0], Jc["T_ps"][:, 0]], axis=1)
    E2p = np.stack([Jc["T_ce"][:, 4], Jc["T_pe"][:, 4], Jc["T_ps"][:, 4]], axis=1)
    # ---- 3. current-collector chains -------------------------------------------------------------------
    ch = g.cha | g.chz
    cin = np.where(g.cha, Jc["Tx_L"], np.where(g.chz, Jc["Tx_U"], 0.0))
    cout = np.where(g.cha, Jc["Tx_U"], np.where(g.chz, Jc["Tx_L"], 0.0))
    di = np.where(ch, Jc["Tx_D"] - cj, 1.0)
    cpred = np.where(g.cha, np.maximum(L - 1, 0), np.minimum(L + 1, 31))
    csucc = np.where(g.cha, np.minimum(L + 1, 31), np.maximum(L - 1, 0))
    cop = shfl(cout, cpred)
    pv = di.copy(); mm = np.zeros(32)
    for _ in range(max(g.Na, g.Nz) - 1):
        pp = shfl(pv, cpred)
        mm = cin / pp
        pv = di - mm * cop
    Fa.chm, Fa.chip, Fa.chup, Fa.cpred, Fa.csucc = mm, 1.0 / pv, cout, cpred, csucc
    Fa.is_tail = (g.cha & (L == g.Na - 1)) | (g.chz & (L == Nx - g.Nz))
    tail = np.where(L == 0, g.Na - 1, Nx - g.Nz)
    hm = np.where((L == 0) | (L == Nx - 1), Jc["T_chain"] / shfl(pv, tail), 0.0)
    D[:, 3, 3] -= hm * shfl(cout, tail)
    Fa.hm = hm
    # ---- 4. twisted block-Thomas -----------------------------------------------------------------------
    left = L < mid; right = (L > mid) & (L < Nx)
    pred = np.where(left, np.maximum(L - 1, 0), np.where(right, np.minimum(L + 1, Nx - 1), np.where(L == mid, L - 1, L)))
    succ = np.where(left, L + 1, np.where(right, L - 1, L))
    pred2 = np.where(left, np.maximum(L - 2, 0), np.where(right, np.minimum(L + 2, Nx - 1), L))
    succ2 = np.where(left, np.minimum(L + 2, Nx - 1), np.where(right, np.maximum(L - 2, 0), L))
    pos = np.where(left, L, np.where(right, Nx - 1 - L, -1))
    is_mid = L == mid
    n_in = max(mid - 1, Nx - mid - 2); n_out = max(mid, Nx - 1 - mid)
    Cin = np.where(right[:, None, None], U9, L9)
    Cout = np.where(right[:, None, None], L9, U9)
    Ein = np.where(right[:, None], E2p, E2m)
    Eout = np.where(right[:, None], E2m, E2p)
    has_pred = pred != L
    Cin[~has_pred] = 0.0
    hasF = np.any(Ein != 0.0, axis=1) & (pos >= 2)
    posA, posB = g.Np - 1, g.Nn - 1
    assert posA - 1 < n_in and posB - 1 < n_in and posA >= 4 and posB >= 4
    Cp = shfl(Cout, pred)
    Di = inv4(D)
    Wm = np.zeros((32, 4, 4)); Fr = np.zeros((32, 4)); Xo = np.zeros((32, 4, 4)); Xp = np.zeros((32, 4, 4))
    for it in range(n_in + 1):
        G = shfl(Di, pred)
        if it == posA - 1 or it == posB - 1:
            mine = (pos == it + 1) & hasF
            G2 = shfl(Di, pred2)
            Cq = shfl(Cout, pred2)
            Fnew = np.einsum("lk,lkc->lc", Ein, G2[:, :3, :])
            Fr = np.where(mine[:, None], Fnew, Fr)
            Cin[:, 3, :] -= np.where(mine[:, None], np.einsum("lk,lkc->lc", Fnew, Cq), 0.0)
        Wm = np.einsum("lrk,lkc->lrc", Cin, G)
        if it == 1:
            D -= np.where((pos == 2)[:, None, None], np.einsum("lrk,lkc->lrc", Wm, Xp), 0.0)
        Dp = D - np.einsum("lrk,lkc->lrc", Wm, Cp)
        Di = np.where(has_pred[:, None, None], inv4(Dp), Di)
        if it == 0:
            Eop = shfl(Eout, pred)
            Xo[:, :, :3] = np.where((pos == 1)[:, None, None], -Wm[:, :, 3][:, :, None] * Eop[:, None, :], 0.0)
            Xp = shfl(Xo, pred)
    Pm = np.einsum("lrk,lkc->lrc", Di, Cout + Xo)
    # meeting node
    G = shfl(Di, np.full(32, mid + 1)); q9 = shfl(L9, np.full(32, mid + 1))
    Wr = np.einsum("lrk,lkc->lrc", U9, G)
    Dp = D - np.einsum("lrk,lkc->lrc", Wm, Cp) - np.einsum("lrk,lkc->lrc", Wr, q9)
    Di = np.where(is_mid[:, None, None], inv4(Dp), Di)
    Pm = np.where(is_mid[:, None, None], Wr, Pm)
    Fa.Dinv, Fa.Wm, Fa.Pm, Fa.Fr, Fa.Eo = Di, Wm, Pm, Fr, Eout
    Fa.pred, Fa.succ, Fa.pred2, Fa.succ2, Fa.pos, Fa.is_mid, Fa.hasF = pred, succ, pred2, succ2, pos, is_mid, hasF
    Fa.n_in, Fa.n_out, Fa.posA, Fa.posB = n_in, n_out, posA, posB
    # ---- 5. border ------------------------------------------------------------------------------------
    zb = np.zeros((32, 4)); zb[:, 2] = Jc["ps_I"]
    Fa.z, Fa.zx = core_solve(g, Fa, zb, Jc["Tx_I"].copy())
    Fa.g_ps0, Fa.g_psN = Jc["g_ps0"], Jc["g_psN"]
    Fa.schur_inv = 1.0 / (Jc["g_I"] - Fa.g_ps0 * Fa.z[0, 2] - Fa.g_psN * Fa.z[Nx - 1, 2])
    return Fa


def solve(g, Fa, rhs):
    """rhs, solution: vectors in the reference layout."""
    def take(idx):
        return np.where(idx >= 0, rhs[np.where(idx >= 0, idx, 0)], 0.0)
    gce, gpe, gps, gT, gj, gTx = take(g.i_ce), take(g.i_pe), take(g.i_ps), take(g.i_T), take(g.i_j), take(g.i_Tx)
    gcs = take(g.i_cs)
    gI = rhs[g.off_I]
    w0 = np.einsum("ic,lc->li", EVI, gcs)
    s9 = np.einsum("i,li,li->l", EV[NR - 1], Fa.pd, w0)
    q0 = np.where(g.elec, (gj - Fa.jcs * s9) * Fa.inv_den, 0.0)
    rb = np.stack([gce - Fa.sj[:, 0] * q0, gpe - Fa.sj[:, 1] * q0, gps - Fa.sj[:, 2] * q0,
                   gT - Fa.sj[:, 3] * q0 - Fa.tcs * s9], axis=1)
    rb[~g.act] = 0.0
    u, ux = core_solve(g, Fa, rb, gTx)
    dI = (gI - Fa.g_ps0 * u[0, 2] - Fa.g_psN * u[g.Nx - 1, 2]) * Fa.schur_inv
    u = u - Fa.z * dI
    ux = ux - Fa.zx * dI
    dj = np.where(g.elec, q0 + np.einsum("lc,lc->l", Fa.q, u), 0.0)
    v = Fa.pd * (w0 - EVI[None, :, NR - 1] * (Fa.csj * dj)[:, None] - Fa.wT * u[:, 3][:, None])
    cs = np.einsum("ri,li->lr", EV, v)
    x = np.zeros_like(rhs)

    def put(idx, val):
        ok = idx >= 0
        x[idx[ok]] = val[ok]
    put(g.i_ce, u[:, 0]); put(g.i_pe, u[:, 1]); put(g.i_ps, u[:, 2]); put(g.i_T, u[:, 3]); put(g.i_j, dj)
    put(g.i_Tx, ux); put(g.i_cs, cs)
    x[g.off_I] = dI
    return x
