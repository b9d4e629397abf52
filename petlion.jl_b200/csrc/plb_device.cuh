// plb_device.cuh -- warp-per-system device code for the PETLION DFN model (sm_100a, FP64).
//
// One warp owns one ~300-equation DAE system; lane x owns x-node x of the 1-D finite-volume grid
// (cathode | separator | anode, N_p+N_s+N_n <= 32) together with that node's particle.
// Neighbour coupling goes through warp shuffles; nothing here touches tensor cores (there is no
// dense contraction in this path).
//
// What this file replaces in the reference (/root/reference/src):
//   residual  R_full  = f_diff! + f_alg! + scalar_residual!   physics_equations/scalar_residual.jl:558-583
//   Jacobian  J_full  = J_y!    + scalar_jacobian!             physics_equations/scalar_residual.jl:588-602
//   physics behind them: physics_equations/residuals.jl:6-106 (c_e), 128-180 (c_s Fickian FD),
//     491-517 (j), 554-654 (Phi_e), 656-703 (Phi_s); auxiliary_states_and_coefficients.jl:6-52
//   KLU factor/solve inside IDA and newtons_method!            model_evaluation.jl:265-271, 417-452
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "laws_generated.cuh"

namespace plb {

constexpr double kF = 96485.3321233;     // const_Faradays,  structures.jl:10
constexpr double kR = 8.31446261815324;  // const_Ideal_Gas, structures.jl:11
constexpr double kTref = 298.15;
constexpr int NR = laws::NR;
constexpr unsigned FULL = 0xffffffffu;

enum { CHEM_LCO = 0, CHEM_NMC = 1 };
enum { METHOD_I = 0, METHOD_V = 1, METHOD_P = 2 };

// canonical parameter fields used by the isothermal model (ASCII names of the reference keys)
enum ThetaField {
    TF_D_n, TF_D_p, TF_D_s, TF_D_sn, TF_D_sp, TF_Ea_D_sn, TF_Ea_D_sp, TF_Ea_k_n, TF_Ea_k_p,
    TF_Rp_n, TF_Rp_p, TF_T0, TF_brugg_n, TF_brugg_p, TF_brugg_s, TF_c_e0, TF_c_max_n, TF_c_max_p,
    TF_k_n, TF_k_p, TF_l_n, TF_l_p, TF_l_s, TF_t_plus, TF_theta_max_n, TF_theta_max_p,
    TF_theta_min_n, TF_theta_min_p, TF_sigma_n, TF_sigma_p, TF_eps_fn, TF_eps_fp, TF_eps_n,
    TF_eps_p, TF_eps_s, TF_COUNT
};

// model descriptor (by value into kernels)
struct ModelDesc {
    int Np, Ns, Nn, Nx, Ne;      // nodes per section, Nx = Np+Ns+Nn <= 32, Ne = Np+Nn
    int chem;                    // CHEM_*
    int ntheta;                  // length of one theta row (reference order, used keys only)
    int theta_stride;            // row stride in doubles
    // reference layout offsets (external.jl:275-365): c_e | c_s (particle-major) | j | Phi_e | Phi_s | I
    int off_cs, off_j, off_pe, off_ps, off_I, N_diff, N_tot;
    int8_t slot[TF_COUNT];       // theta field -> position in the row (-1: not a key of this variant)
};

// ------------------------------------------------------------------------------------------------
// per-warp shared-memory constants derived from theta once per system
// ------------------------------------------------------------------------------------------------
enum SecField {
    SC_h, SC_inv_h, SC_inv_por, SC_pb, SC_Dlin, SC_src_ce, SC_hFa, SC_psf, SC_kap, SC_inv_Rp,
    SC_Rp_Ds, SC_k2, SC_cmax, SC_inv_cmax, SC_COUNT
};
enum GlobField { GC_T, GC_xcoef, GC_Kc, GC_I1C, GC_psI_p, GC_psI_n, GC_dUdT_on, GC_COUNT };

struct WarpConst {
    double sec[SC_COUNT][4];   // [field][section p,s,n,pad]
    double g[GC_COUNT + 1];
    double dinv[32];           // 1 / (centre distance across face x|x+1)
    double beta[32];           // harmonic-mean weight of face x|x+1
    double theta[TF_COUNT + 1];
};

struct LaneRole {
    int x;        // node index == lane
    int sec;      // 0 p, 1 s, 2 n, 3 inactive
    int e;        // electrode index (p: x, n: x-Ns), -1 otherwise
    bool act, elec, first_e, last_e;   // first/last node of its electrode
};

__device__ __forceinline__ LaneRole make_role(const ModelDesc& m, int lane) {
    LaneRole r;
    r.x = lane;
    r.act = lane < m.Nx;
    r.sec = lane < m.Np ? 0 : (lane < m.Np + m.Ns ? 1 : (lane < m.Nx ? 2 : 3));
    r.elec = (r.sec == 0) || (r.sec == 2);
    r.e = r.sec == 0 ? lane : (r.sec == 2 ? lane - m.Ns : -1);
    r.first_e = (r.sec == 0 && lane == 0) || (r.sec == 2 && lane == m.Np + m.Ns);
    r.last_e = (r.sec == 0 && lane == m.Np - 1) || (r.sec == 2 && lane == m.Nx - 1);
    return r;
}

// one node's unknowns held in registers
struct LaneVec {
    double ce, cs[NR], j, pe, ps;
};

__device__ __forceinline__ double shfl_dn(double v) { return __shfl_down_sync(FULL, v, 1); }
__device__ __forceinline__ double shfl_up(double v) { return __shfl_up_sync(FULL, v, 1); }
__device__ __forceinline__ double shfl_from(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// calc_I1C -- auxiliary_states_and_coefficients.jl:631-647
__device__ __forceinline__ double calc_I1C(const double* th) {
    const double eps_sp = 1.0 - (th[TF_eps_fp] + th[TF_eps_p]);
    const double eps_sn = 1.0 - (th[TF_eps_fn] + th[TF_eps_n]);
    const double qp = eps_sp * th[TF_l_p] * th[TF_c_max_p] * (th[TF_theta_min_p] - th[TF_theta_max_p]);
    const double qn = eps_sn * th[TF_l_n] * th[TF_c_max_n] * (th[TF_theta_max_n] - th[TF_theta_min_n]);
    return (kF / 3600.0) * fmin(qp, qn);
}

// Fill WarpConst from one theta row (global or shared).  All lanes participate.
// build_auxiliary_states! -- auxiliary_states_and_coefficients.jl:6-52 (parameter-only parts)
__device__ __forceinline__ void setup_consts(const ModelDesc& m, const double* __restrict__ theta_row,
                                             WarpConst& C, int lane) {
    for (int f = lane; f < TF_COUNT; f += 32) {
        const int s = m.slot[f];
        C.theta[f] = s >= 0 ? theta_row[s] : 0.0;
    }
    __syncwarp();
    const double* th = C.theta;
    if (lane < 3) {
        const int s = lane;
        const double eps_f = s == 0 ? th[TF_eps_fp] : th[TF_eps_fn];
        const double eps_e = s == 0 ? th[TF_eps_p] : th[TF_eps_n];
        const double eps_act = 1.0 - (eps_f + eps_e);                                  // active_material
        const double por = s == 1 ? th[TF_eps_s] : 1.0 - (eps_f + eps_act);            // build_eps!
        const double brg = s == 0 ? th[TF_brugg_p] : (s == 1 ? th[TF_brugg_s] : th[TF_brugg_n]);
        const double l = s == 0 ? th[TF_l_p] : (s == 1 ? th[TF_l_s] : th[TF_l_n]);
        const int n = s == 0 ? m.Np : (s == 1 ? m.Ns : m.Nn);
        const double h = (1.0 / n) * l;
        const double Rp = s == 0 ? th[TF_Rp_p] : th[TF_Rp_n];
        const double a = s == 1 ? 0.0 : 3 * eps_act / Rp;                               // build_a!
        const double sig = (s == 0 ? th[TF_sigma_p] : th[TF_sigma_n]) * eps_act;       // build_sigma_eff
        const double Dl = s == 0 ? th[TF_D_p] : (s == 1 ? th[TF_D_s] : th[TF_D_n]);
        const double pb = pow(por, brg);
        const double T = th[TF_T0];
        const bool Tref = (T == kTref);   // temperature_switch, custom_functions.jl:1
        double Ds = s == 0 ? th[TF_D_sp] : th[TF_D_sn];
        double k = s == 0 ? th[TF_k_p] : th[TF_k_n];
        if (!Tref && s != 1) {            // D_s_eff / rxn_rate Arrhenius, custom_functions.jl:16-57
            const double EaD = s == 0 ? th[TF_Ea_D_sp] : th[TF_Ea_D_sn];
            const double Eak = s == 0 ? th[TF_Ea_k_p] : th[TF_Ea_k_n];
            Ds = Ds * exp(-EaD / kR * (1.0 / T - 1.0 / kTref));
            k = k * exp(-(Eak / kR) * (1.0 / T - 1.0 / kTref));
        }
        const double cmax = s == 0 ? th[TF_c_max_p] : th[TF_c_max_n];
        C.sec[SC_h][s] = h;
        C.sec[SC_inv_h][s] = 1.0 / h;
        C.sec[SC_inv_por][s] = 1.0 / por;
        C.sec[SC_pb][s] = pb;
        C.sec[SC_Dlin][s] = Dl * pb;
        C.sec[SC_src_ce][s] = (1 - th[TF_t_plus]) * a;
        C.sec[SC_hFa][s] = h * kF * a;
        C.sec[SC_psf][s] = s == 1 ? 0.0 : h * h * a * kF / sig;
        C.sec[SC_kap][s] = s == 1 ? 0.0 : Ds / (Rp * Rp);
        C.sec[SC_inv_Rp][s] = s == 1 ? 0.0 : 1.0 / Rp;
        C.sec[SC_Rp_Ds][s] = s == 1 ? 0.0 : Rp / Ds;
        C.sec[SC_k2][s] = s == 1 ? 0.0 : 2.0 * k;
        C.sec[SC_cmax][s] = cmax;
        C.sec[SC_inv_cmax][s] = 1.0 / cmax;
        if (s == 0) {
            const double I1C = calc_I1C(th);
            C.g[GC_T] = T;
            C.g[GC_xcoef] = 0.5 * kF / (kR * T);
            C.g[GC_Kc] = 2 * kR * (1 - th[TF_t_plus]) / kF;
            C.g[GC_I1C] = I1C;
            C.g[GC_psI_p] = I1C * h / sig;     // d res_Phi_s[first p] / dI   (residuals.jl:679)
            C.g[GC_dUdT_on] = (m.chem == CHEM_LCO && !Tref) ? 1.0 : 0.0;
        }
        if (s == 2) C.g[GC_psI_n] = -calc_I1C(th) * h / sig;   // d res_Phi_s[last n] / dI (residuals.jl:680)
    }
    __syncwarp();
    {
        // face x | x+1 geometry: numerical_tools.jl:106-215
        const int x = lane;
        const int s0 = x < m.Np ? 0 : (x < m.Np + m.Ns ? 1 : 2);
        const int x1 = x + 1;
        const int s1 = x1 < m.Np ? 0 : (x1 < m.Np + m.Ns ? 1 : 2);
        const double h0 = C.sec[SC_h][s0], h1 = C.sec[SC_h][s1];
        double d, b;
        if (s0 == s1) { d = h0; b = 0.5; }
        else { d = h0 / 2 + h1 / 2; b = (h0 / 2) / (h1 / 2 + h0 / 2); }
        C.dinv[lane] = (x < m.Nx - 1) ? 1.0 / d : 0.0;
        C.beta[lane] = b;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// residual + Jacobian coefficients of one node (lane)
// ------------------------------------------------------------------------------------------------
struct LaneJac {
    // c_e row: d/d ce[x-1], ce[x] (without the -cj term), ce[x+1], d/dj
    double ceL, ceD, ceU, ce_j;
    // j row: d/d cs_surf, ce, pe, ps   (d/dj = -1)
    double j_cs, j_ce, j_pe, j_ps;
    // Phi_e row: d/d pe[x-1..x+1], d/d ce[x-1..x+1], d/dj
    double peL, peD, peU, pcL, pcD, pcU, pe_j;
    // Phi_s row: tridiagonal constants, d/dj, d/dI
    double psL, psD, psU, ps_j, ps_I;
    // particle: block = kap*MC - cj*I ; surface-row d/dj
    double kap, cs_j;
};

// Control row (scalar_residual.jl:167-202): residual and its three possible Jacobian entries
struct CtrlRow {
    double res, g_ps0, g_psN, g_I;
};

template <int CHEM, bool WITH_JAC>
__device__ __forceinline__ void lane_eval(const ModelDesc& m, const WarpConst& C, const LaneRole& ro,
                                          const LaneVec& y, const LaneVec& yp, double Iapp,
                                          int method, double value, LaneVec& res, CtrlRow& ctrl,
                                          LaneJac& J) {
    const int s = ro.sec < 3 ? ro.sec : 1;
    const double T = C.g[GC_T];
    // ---- node-local electrolyte properties: build_K_eff!/build_D_eff! (:302-328) -----------------
    double K = 0.0, dK = 0.0, D = 0.0, dD = 0.0;
    const double ce = ro.act ? y.ce : 1000.0;
    {
        const double pb = C.sec[SC_pb][s];
        laws::K_eff(ce, T, K, dK);
        K *= pb; dK *= pb;
        if (CHEM == CHEM_LCO) { D = C.sec[SC_Dlin][s]; dD = 0.0; }   // D_eff_linear
        else { laws::D_eff_nl(ce, T, D, dD); D *= pb; dD *= pb; }
    }
    // ---- face x|x+1 (owned by lane x): harmonic means and fluxes ---------------------------------
    const double ceR = shfl_dn(ce), peR = shfl_dn(y.pe), KR = shfl_dn(K), dKR = shfl_dn(dK),
                 DR = shfl_dn(D), dDR = shfl_dn(dD);
    const bool has_face = ro.x < m.Nx - 1;
    const double b = C.beta[ro.x], dinv = C.dinv[ro.x];
    double Nf = 0.0, Q = 0.0;
    double dN_cL = 0.0, dN_cR = 0.0, dQ_cL = 0.0, dQ_cR = 0.0, wK = 0.0;
    if (has_face) {
        const double denK = b * KR + (1.0 - b) * K;
        const double Khat = K * KR / denK;                          // interpolate_electrolyte_grid
        const double denD = b * DR + (1.0 - b) * D;
        const double Dhat = D * DR / denD;
        const double denc = b * ceR + (1.0 - b) * ce;
        const double cbar = ce * ceR / denc;                        // interpolate_electrolyte_concentration
        const double Tbar = T * T / (b * T + (1.0 - b) * T);        // interpolate_temperature
        const double dc = (ceR - ce) * dinv;                        // ..._concetration_fluxes
        const double G = Khat * Tbar * dc / cbar;                   // prod_tot, residuals.jl:631-635
        wK = Khat * dinv;
        Q = wK * (peR - y.pe) - C.g[GC_Kc] * G;
        Nf = Dhat * dinv * (ceR - ce);
        if (WITH_JAC) {
            const double iK2 = 1.0 / (denK * denK), iD2 = 1.0 / (denD * denD), ic2 = 1.0 / (denc * denc);
            const double dKh_cL = b * KR * KR * iK2 * dK, dKh_cR = (1.0 - b) * K * K * iK2 * dKR;
            const double dDh_cL = b * DR * DR * iD2 * dD, dDh_cR = (1.0 - b) * D * D * iD2 * dDR;
            const double dcb_cL = b * ceR * ceR * ic2, dcb_cR = (1.0 - b) * ce * ce * ic2;
            const double icb = 1.0 / cbar;
            const double dG_cL = Tbar * icb * (dKh_cL * dc - Khat * dinv - Khat * dc * dcb_cL * icb);
            const double dG_cR = Tbar * icb * (dKh_cR * dc + Khat * dinv - Khat * dc * dcb_cR * icb);
            dQ_cL = dKh_cL * dinv * (peR - y.pe) - C.g[GC_Kc] * dG_cL;
            dQ_cR = dKh_cR * dinv * (peR - y.pe) - C.g[GC_Kc] * dG_cR;
            dN_cL = -Dhat * dinv + dDh_cL * dc;
            dN_cR = Dhat * dinv + dDh_cR * dc;
        }
    }
    // left face x-1|x comes from lane x-1
    double NfL = shfl_up(Nf), QL = shfl_up(Q), wKL = shfl_up(wK);
    double dNL_cL = 0.0, dNL_cR = 0.0, dQL_cL = 0.0, dQL_cR = 0.0;
    if (WITH_JAC) { dNL_cL = shfl_up(dN_cL); dNL_cR = shfl_up(dN_cR); dQL_cL = shfl_up(dQ_cL); dQL_cR = shfl_up(dQ_cR); }
    if (ro.x == 0) { NfL = 0.0; QL = 0.0; wKL = 0.0; dNL_cL = dNL_cR = dQL_cL = dQL_cR = 0.0; }

    // ---- electrode-node quantities ------------------------------------------------------------------
    double jtot = 0.0, jcalc = 0.0;
    double dj_cs = 0.0, dj_ce = 0.0, dj_eta = 0.0;
    if (ro.elec) {
        jtot = y.j;                                                          // build_j_total!
        const double cs_s = y.cs[NR - 1];                                    // build_c_s_star!
        const double th = cs_s * C.sec[SC_inv_cmax][s];
        double U, dU, dUdT = 0.0, ddUdT = 0.0;
        if (CHEM == CHEM_LCO) {
            if (ro.sec == 0) laws::OCV_LCO(th, U, dU, dUdT, ddUdT);
            else {
                // sqrt_ReLU branches (custom_functions.jl:143, 210): physical range th > 1e-4
                const double sq = sqrt(fmax(th, 1e-4));
                laws::OCV_LiC6(th, sq, U, dU, dUdT, ddUdT);
            }
            if (C.g[GC_dUdT_on] != 0.0) { U += dUdT * (T - kTref); dU += ddUdT * (T - kTref); }
        } else {
            if (ro.sec == 0) laws::OCV_NMC(th, U, dU);
            else laws::OCV_LiC6_NMC(th, U, dU);
        }
        const double eta = y.ps - y.pe - U;                                  // build_eta!
        // rxn_BV, custom_functions.jl:212-231
        const double cmax = C.sec[SC_cmax][s];
        const double arg = ce * cs_s * (cmax - cs_s);
        const double sq = arg > 0.0 ? sqrt(arg) : 0.0;                       // sqrt_ReLU
        const double xx = C.g[GC_xcoef] * eta;
        const double em = expm1(xx);
        const double sh = 0.5 * (em + em / (em + 1.0));                      // sinh(xx)
        const double k2 = C.sec[SC_k2][s];
        jcalc = k2 * sq * sh;
        if (WITH_JAC) {
            const double ch = sh + 1.0 / (em + 1.0);                         // cosh = sinh + exp(-x)
            const double isq = arg > 0.0 ? 0.5 / sq : 0.0;
            dj_eta = k2 * sq * ch * C.g[GC_xcoef];
            dj_ce = k2 * sh * isq * cs_s * (cmax - cs_s);
            dj_cs = k2 * sh * isq * ce * (cmax - 2.0 * cs_s) - dj_eta * dU * C.sec[SC_inv_cmax][s];
        }
    }

    // ---- residuals_c_e!, residuals.jl:6-106 ---------------------------------------------------------
    const double inv_h = C.sec[SC_inv_h][s], inv_por = C.sec[SC_inv_por][s];
    res.ce = ((Nf - NfL) * inv_h + C.sec[SC_src_ce][s] * jtot) * inv_por - yp.ce;
    // ---- residuals_Phi_e!, residuals.jl:554-654 -----------------------------------------------------
    const bool last = ro.x == m.Nx - 1;
    res.pe = last ? y.pe : (QL - Q - C.sec[SC_hFa][s] * jtot);
    // ---- residuals_c_s_avg! (Fickian FD), residuals.jl:128-180 --------------------------------------
    // ---- residuals_j!, residuals.jl:491-517 ; residuals_Phi_s!, residuals.jl:656-703 ----------------
    const double psL = shfl_up(y.ps), psR = shfl_dn(y.ps);
    if (ro.elec) {
        const double kap = C.sec[SC_kap][s];
        const double d1bc = -y.j * C.sec[SC_Rp_Ds][s];
#pragma unroll
        for (int r = 0; r < NR; r++) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < NR; c++)
                if (laws::mc_mask(r) & (1u << c)) acc = fma(laws::MC[r][c], y.cs[c], acc);
            if (r == NR - 1) acc = fma(laws::BJ, d1bc, acc);
            res.cs[r] = kap * acc - yp.cs[r];
        }
        res.j = jcalc - y.j;
        double f = C.sec[SC_psf][s] * jtot;
        double acc = -(ro.first_e || ro.last_e ? 1.0 : 2.0) * y.ps;
        if (!ro.first_e) acc += psL;
        if (!ro.last_e) acc += psR;
        if (ro.sec == 0 && ro.first_e) f -= C.g[GC_psI_p] * Iapp;
        if (ro.sec == 2 && ro.last_e) f -= C.g[GC_psI_n] * Iapp;      // psI_n carries the sign
        res.ps = acc - f;
    } else {
#pragma unroll
        for (int r = 0; r < NR; r++) res.cs[r] = 0.0;
        res.j = 0.0;
        res.ps = 0.0;
    }
    // ---- control row: scalar_residual!, scalar_residual.jl:167; calc_V/P :86-87 ----------------------
    {
        const double ps0 = shfl_from(y.ps, 0), psN = shfl_from(y.ps, m.Nx - 1);
        const double V = ps0 - psN;
        if (method == METHOD_I) { ctrl.res = Iapp - value; ctrl.g_ps0 = 0.0; ctrl.g_psN = 0.0; ctrl.g_I = 1.0; }
        else if (method == METHOD_V) { ctrl.res = V - value; ctrl.g_ps0 = 1.0; ctrl.g_psN = -1.0; ctrl.g_I = 0.0; }
        else {
            const double I1C = C.g[GC_I1C];
            ctrl.res = Iapp * I1C * V - value;
            ctrl.g_ps0 = Iapp * I1C; ctrl.g_psN = -Iapp * I1C; ctrl.g_I = V * I1C;
        }
    }
    if (WITH_JAC) {
        const double ihp = inv_h * inv_por;
        J.ceL = -dNL_cL * ihp;
        J.ceD = (dN_cL - dNL_cR) * ihp;
        J.ceU = dN_cR * ihp;
        J.ce_j = ro.elec ? C.sec[SC_src_ce][s] * inv_por : 0.0;
        J.j_cs = dj_cs; J.j_ce = dj_ce; J.j_pe = -dj_eta; J.j_ps = dj_eta;
        if (last) {
            J.peL = 0.0; J.peD = 1.0; J.peU = 0.0; J.pcL = J.pcD = J.pcU = 0.0; J.pe_j = 0.0;
        } else {
            J.peL = -wKL; J.peD = wKL + wK; J.peU = -wK;
            J.pcL = dQL_cL; J.pcD = dQL_cR - dQ_cL; J.pcU = -dQ_cR;
            J.pe_j = ro.elec ? -C.sec[SC_hFa][s] : 0.0;
        }
        J.psL = (ro.elec && !ro.first_e) ? 1.0 : 0.0;
        J.psU = (ro.elec && !ro.last_e) ? 1.0 : 0.0;
        J.psD = ro.elec ? -((ro.first_e || ro.last_e) ? 1.0 : 2.0) : 1.0;
        J.ps_j = ro.elec ? -C.sec[SC_psf][s] : 0.0;
        J.ps_I = (ro.sec == 0 && ro.first_e) ? C.g[GC_psI_p] : ((ro.sec == 2 && ro.last_e) ? C.g[GC_psI_n] : 0.0);
        J.kap = C.sec[SC_kap][s];
        J.cs_j = ro.elec ? -laws::BJ * C.sec[SC_inv_Rp][s] : 0.0;
    }
}

// out-of-line copy for the fused integrator (keeps its code inside the instruction cache)
template <int CHEM, bool WITH_JAC>
__device__ __noinline__ void lane_eval_ni(const ModelDesc& m, const WarpConst& C, const LaneRole& ro,
                                          const LaneVec& y, const LaneVec& yp, double Iapp, int method,
                                          double value, LaneVec& res, CtrlRow& ctrl, LaneJac& J) {
    lane_eval<CHEM, WITH_JAC>(m, C, ro, y, yp, Iapp, method, value, res, ctrl, J);
}

// ------------------------------------------------------------------------------------------------
// structured Newton-matrix factorisation / solve
//   1. particle block (kap*MC - cj I) is identical for every particle of an electrode -> explicit
//      10x10 inverse per electrode (Gauss-Jordan, no pivoting needed: growth <= 1.4 measured)
//   2. eliminate c_s (through its surface value) and j node-locally
//   3. 3x3-block tridiagonal system in (c_e, Phi_e, Phi_s) over the nodes: block Thomas along lanes
//   4. the applied-current unknown I is a border (Schur complement), which also covers the
//      zero-diagonal control row of voltage/power control (scalar_residual.jl:184-197)
// ------------------------------------------------------------------------------------------------
struct WarpFactor {
    double Sinv[NR * NR][2];   // [r*NR+c][electrode 0=p,1=n]
    double vb[NR][2];          // Sinv * b (b = surface-row j coupling)
    // per-lane data [field][lane]
    double Dinv[9][32];        // inverse of the pivoted 3x3 diagonal block D'_x
    double Wm[9][32];          // W_x = L_x * Dinv_{x-1}     (forward sweep:  y_x = r_x - W_x y_{x-1})
    double Pm[9][32];          // P_x = Dinv_x * U_x         (backward sweep: u_x = Dinv_x y_x - P_x u_{x+1})
    double z[3][32];           // T^{-1} e_I  (border column)
    double q[4][32];           // j elimination: q_ce, q_pe, q_ps, inv_den
    double jcs[32];            // a_cs (j row coefficient of the surface concentration)
    double sj[3][32];          // d(row)/dj for rows ce, pe, ps
    double schur_inv;          // 1/(g_I - g_ps0*z_ps[0] - g_psN*z_ps[N-1])
    double g_ps0, g_psN;
    double pad;
};

// 3x3 inverse by the adjugate (forward error ~ cond * eps, invariant under row/column scaling)
__device__ __forceinline__ void inv3x3(const double* a, double* o) {
    const double c00 = a[4] * a[8] - a[5] * a[7];
    const double c01 = a[5] * a[6] - a[3] * a[8];
    const double c02 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    const double id = 1.0 / det;
    o[0] = c00 * id;
    o[3] = c01 * id;
    o[6] = c02 * id;
    o[1] = (a[2] * a[7] - a[1] * a[8]) * id;
    o[4] = (a[0] * a[8] - a[2] * a[6]) * id;
    o[7] = (a[1] * a[6] - a[0] * a[7]) * id;
    o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// Block-Thomas as a "twisted" (two-sided) elimination: nodes 0..mid-1 are eliminated left-to-right,
// nodes Nx-1..mid+1 right-to-left, both chains meet at node `mid`, and the back-substitution runs
// outward from `mid` in both directions at once.  That halves the serial depth (15 instead of 29
// dependent block steps per sweep for 30 nodes).  Each sweep is written as a select-free fixed-point
// iteration along the lanes: after k iterations the k-th node of each chain holds its final value and
// never changes again, so the iteration reproduces the sequential recurrences exactly.
struct LaneChain {
    int pred, succ;      // lane this node takes its value from: inward sweep (pred), outward sweep (succ)
    bool is_mid;
    int n_in, n_out;     // iterations needed by the inward / outward sweeps
};
__device__ __forceinline__ LaneChain make_chain(int Nx, int lane) {
    LaneChain c;
    const int mid = Nx / 2;
    c.is_mid = lane == mid;
    const bool left = lane < mid, right = lane > mid && lane < Nx;
    c.pred = left ? (lane > 0 ? lane - 1 : lane) : (right ? (lane < Nx - 1 ? lane + 1 : lane) : (lane == mid ? lane - 1 : lane));
    c.succ = left ? lane + 1 : (right ? lane - 1 : lane);
    const int a = mid - 1, b = Nx - mid - 2;
    c.n_in = a > b ? a : b;
    c.n_out = mid > Nx - 1 - mid ? mid : Nx - 1 - mid;
    return c;
}
//   inward :  y_x = r_x - W_x y_pred(x)                    (chain heads: W = 0)
//             y_mid = r_mid - Wl y_{mid-1} - Wr y_{mid+1}
//   outward:  u_mid = Dinv_mid y_mid ;  u_x = Dinv_x y_x - P_x u_succ(x)
// Wm: W_x (for mid: Wl);  Pm: P_x (for mid: Wr, which has no outward update of its own)
__device__ __forceinline__ void thomas_sweeps(int Nx, const LaneChain& ch, const double* Wm, const double* Pm,
                                              const double* Di, const double* r, double* u) {
    double y0 = r[0], y1 = r[1], y2 = r[2];
#pragma unroll 1
    for (int it = 0; it < ch.n_in; it++) {
        const double a0 = shfl_from(y0, ch.pred), a1 = shfl_from(y1, ch.pred), a2 = shfl_from(y2, ch.pred);
        y0 = r[0] - (Wm[0] * a0 + Wm[1] * a1 + Wm[2] * a2);
        y1 = r[1] - (Wm[3] * a0 + Wm[4] * a1 + Wm[5] * a2);
        y2 = r[2] - (Wm[6] * a0 + Wm[7] * a1 + Wm[8] * a2);
    }
    {   // the meeting node takes both neighbours
        const int mid = Nx / 2;
        const double a0 = shfl_from(y0, mid - 1), a1 = shfl_from(y1, mid - 1), a2 = shfl_from(y2, mid - 1);
        const double b0 = shfl_from(y0, mid + 1), b1 = shfl_from(y1, mid + 1), b2 = shfl_from(y2, mid + 1);
        if (ch.is_mid) {
            y0 = r[0] - (Wm[0] * a0 + Wm[1] * a1 + Wm[2] * a2) - (Pm[0] * b0 + Pm[1] * b1 + Pm[2] * b2);
            y1 = r[1] - (Wm[3] * a0 + Wm[4] * a1 + Wm[5] * a2) - (Pm[3] * b0 + Pm[4] * b1 + Pm[5] * b2);
            y2 = r[2] - (Wm[6] * a0 + Wm[7] * a1 + Wm[8] * a2) - (Pm[6] * b0 + Pm[7] * b1 + Pm[8] * b2);
        }
    }
    const double c0 = Di[0] * y0 + Di[1] * y1 + Di[2] * y2;
    const double c1 = Di[3] * y0 + Di[4] * y1 + Di[5] * y2;
    const double c2 = Di[6] * y0 + Di[7] * y1 + Di[8] * y2;
    double u0 = c0, u1 = c1, u2 = c2;
    // the meeting node has no outward update: give it a zero P
    const double p0 = ch.is_mid ? 0.0 : Pm[0], p1 = ch.is_mid ? 0.0 : Pm[1], p2 = ch.is_mid ? 0.0 : Pm[2];
    const double p3 = ch.is_mid ? 0.0 : Pm[3], p4 = ch.is_mid ? 0.0 : Pm[4], p5 = ch.is_mid ? 0.0 : Pm[5];
    const double p6 = ch.is_mid ? 0.0 : Pm[6], p7 = ch.is_mid ? 0.0 : Pm[7], p8 = ch.is_mid ? 0.0 : Pm[8];
#pragma unroll 1
    for (int it = 0; it < ch.n_out; it++) {
        const double a0 = shfl_from(u0, ch.succ), a1 = shfl_from(u1, ch.succ), a2 = shfl_from(u2, ch.succ);
        u0 = c0 - (p0 * a0 + p1 * a1 + p2 * a2);
        u1 = c1 - (p3 * a0 + p4 * a1 + p5 * a2);
        u2 = c2 - (p6 * a0 + p7 * a1 + p8 * a2);
    }
    u[0] = u0; u[1] = u1; u[2] = u2;
}

// alg_only: Newton on the algebraic block (newtons_method!, model_evaluation.jl:430-480):
// c_e and c_s are frozen, the differential rows are replaced by identity.
__device__ __forceinline__ void warp_factor_impl(const ModelDesc& m, const LaneRole& ro, const LaneJac& J,
                                                 const CtrlRow& ctrl, double cj, bool alg_only,
                                                 WarpFactor& Fa, int lane) {
    // ---- 1. particle inverses -------------------------------------------------------------------
    if (!alg_only) {
        const int el = lane >> 4;                       // lanes 0-15 -> cathode, 16-31 -> anode
        const int l16 = lane & 15;
        const double kap = shfl_from(J.kap, el == 0 ? 0 : m.Nx - 1);
        for (int k = l16; k < NR * NR; k += 16) {
            const int r = k / NR, c = k - r * NR;
            Fa.Sinv[k][el] = kap * laws::MC[r][c] - (r == c ? cj : 0.0);
        }
        __syncwarp();
#pragma unroll 1
        for (int p = 0; p < NR; p++) {
            const double piv = 1.0 / Fa.Sinv[p * NR + p][el];
            __syncwarp();
            if (l16 < NR && l16 != p) Fa.Sinv[p * NR + l16][el] *= piv;
            __syncwarp();
            for (int k = l16; k < NR * NR; k += 16) {
                const int r = k / NR, c = k - r * NR;
                if (r != p && c != p)
                    Fa.Sinv[k][el] = fma(-Fa.Sinv[r * NR + p][el], Fa.Sinv[p * NR + c][el], Fa.Sinv[k][el]);
            }
            __syncwarp();
            if (l16 < NR) Fa.Sinv[l16 * NR + p][el] = (l16 == p) ? piv : -Fa.Sinv[l16 * NR + p][el] * piv;
            __syncwarp();
        }
        // vb = Sinv * b,  b = cs_j * e_surf
        const double csj = shfl_from(J.cs_j, el == 0 ? 0 : m.Nx - 1);
        if (l16 < NR) Fa.vb[l16][el] = Fa.Sinv[l16 * NR + NR - 1][el] * csj;
        __syncwarp();
    }
    // ---- 2. node-local elimination of c_s and j -------------------------------------------------
    const int el = ro.sec == 2 ? 1 : 0;
    double q_ce = 0.0, q_pe = 0.0, q_ps = 0.0, inv_den = 0.0;
    if (ro.elec) {
        // dcs_surf = p0 + p1*dj with p1 = -(Sinv b)[surf]
        const double p1 = alg_only ? 0.0 : -Fa.vb[NR - 1][el];
        inv_den = 1.0 / (-1.0 + J.j_cs * p1);
        q_ce = alg_only ? 0.0 : -J.j_ce * inv_den;
        q_pe = -J.j_pe * inv_den;
        q_ps = -J.j_ps * inv_den;
    }
    Fa.q[0][lane] = q_ce; Fa.q[1][lane] = q_pe; Fa.q[2][lane] = q_ps; Fa.q[3][lane] = inv_den;
    Fa.jcs[lane] = J.j_cs;
    const double sj0 = alg_only ? 0.0 : J.ce_j, sj1 = J.pe_j, sj2 = J.ps_j;
    Fa.sj[0][lane] = sj0; Fa.sj[1][lane] = sj1; Fa.sj[2][lane] = sj2;
    // diagonal block rows (ce, pe, ps) x cols (ce, pe, ps)
    double Dm[9];
    if (alg_only) { Dm[0] = 1.0; Dm[1] = 0.0; Dm[2] = 0.0; }
    else { Dm[0] = J.ceD - cj + sj0 * q_ce; Dm[1] = sj0 * q_pe; Dm[2] = sj0 * q_ps; }
    Dm[3] = (alg_only ? 0.0 : J.pcD) + sj1 * q_ce; Dm[4] = J.peD + sj1 * q_pe; Dm[5] = sj1 * q_ps;
    Dm[6] = sj2 * q_ce; Dm[7] = sj2 * q_pe; Dm[8] = J.psD + sj2 * q_ps;
    if (!ro.act) { Dm[0] = 1; Dm[1] = 0; Dm[2] = 0; Dm[3] = 0; Dm[4] = 1; Dm[5] = 0; Dm[6] = 0; Dm[7] = 0; Dm[8] = 1; }
    // off-diagonal blocks: L = [[L0,0,0],[L1,L2,0],[0,0,L3]] (entries (ce,ce) (pe,ce) (pe,pe) (ps,ps)), U alike
    double L4[4] = {alg_only ? 0.0 : J.ceL, alg_only ? 0.0 : J.pcL, J.peL, J.psL};
    double U4[4] = {alg_only ? 0.0 : J.ceU, alg_only ? 0.0 : J.pcU, J.peU, J.psU};
    if (!ro.act || ro.x == 0) { L4[0] = L4[1] = L4[2] = L4[3] = 0.0; }
    if (!ro.act || ro.x >= m.Nx - 1) { U4[0] = U4[1] = U4[2] = U4[3] = 0.0; }
    // ---- 3. twisted block-Thomas factorisation as a fixed-point iteration along the lanes ----------
    //   left chain  (x < mid): Cin = L_x, pred = x-1, Cout(pred) = U_{x-1}, Cout(x) = U_x
    //   right chain (x > mid): Cin = U_x, pred = x+1, Cout(pred) = L_{x+1}, Cout(x) = L_x
    //   D'_x = D_x - W_x Cout(pred),  W_x = Cin_x Dinv_pred ;  mid node: both neighbours
    const LaneChain ch = make_chain(m.Nx, lane);
    const int mid = m.Nx / 2;
    const bool rightc = lane > mid && lane < m.Nx;
    double Cin[4], Cout[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { Cin[k] = rightc ? U4[k] : L4[k]; Cout[k] = rightc ? L4[k] : U4[k]; }
    // Cout of the predecessor as seen from this node: left chain needs U_{x-1}, right chain L_{x+1}
    double Cp[4];
    {
        const double u0 = shfl_from(U4[0], ch.pred), u1 = shfl_from(U4[1], ch.pred), u2 = shfl_from(U4[2], ch.pred), u3s = shfl_from(U4[3], ch.pred);
        const double l0 = shfl_from(L4[0], ch.pred), l1 = shfl_from(L4[1], ch.pred), l2 = shfl_from(L4[2], ch.pred), l3 = shfl_from(L4[3], ch.pred);
        Cp[0] = rightc ? l0 : u0; Cp[1] = rightc ? l1 : u1; Cp[2] = rightc ? l2 : u2; Cp[3] = rightc ? l3 : u3s;
    }
    const bool has_pred = ch.pred != lane;
    if (!has_pred) { Cin[0] = Cin[1] = Cin[2] = Cin[3] = 0.0; }
    double Di[9], Wm[9];
    inv3x3(Dm, Di);
#pragma unroll
    for (int k = 0; k < 9; k++) Wm[k] = 0.0;
#pragma unroll 1
    for (int it = 0; it < ch.n_in + 1; it++) {
        double G[9];
#pragma unroll
        for (int k = 0; k < 9; k++) G[k] = shfl_from(Di[k], ch.pred);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            Wm[c] = Cin[0] * G[c];
            Wm[3 + c] = Cin[1] * G[c] + Cin[2] * G[3 + c];
            Wm[6 + c] = Cin[3] * G[6 + c];
        }
        double Dp[9];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            Dp[i * 3 + 0] = Dm[i * 3 + 0] - (Wm[i * 3 + 0] * Cp[0] + Wm[i * 3 + 1] * Cp[1]);
            Dp[i * 3 + 1] = Dm[i * 3 + 1] - Wm[i * 3 + 1] * Cp[2];
            Dp[i * 3 + 2] = Dm[i * 3 + 2] - Wm[i * 3 + 2] * Cp[3];
        }
        if (has_pred) inv3x3(Dp, Di);
    }
    double Pm[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Pm[i * 3 + 0] = Di[i * 3 + 0] * Cout[0] + Di[i * 3 + 1] * Cout[1];
        Pm[i * 3 + 1] = Di[i * 3 + 1] * Cout[2];
        Pm[i * 3 + 2] = Di[i * 3 + 2] * Cout[3];
    }
    {   // meeting node: second neighbour (mid+1, right chain): Wr = U_mid Dinv_{mid+1}, stored in Pm
        double G[9];
#pragma unroll
        for (int k = 0; k < 9; k++) G[k] = shfl_from(Di[k], mid + 1);
        const double q0 = shfl_from(L4[0], mid + 1), q1 = shfl_from(L4[1], mid + 1), q2 = shfl_from(L4[2], mid + 1), q3 = shfl_from(L4[3], mid + 1);
        if (ch.is_mid) {
            double Wr[9];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                Wr[c] = U4[0] * G[c];
                Wr[3 + c] = U4[1] * G[c] + U4[2] * G[3 + c];
                Wr[6 + c] = U4[3] * G[6 + c];
            }
            double Dp[9];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                Dp[i * 3 + 0] = Dm[i * 3 + 0] - (Wm[i * 3 + 0] * Cp[0] + Wm[i * 3 + 1] * Cp[1]) - (Wr[i * 3 + 0] * q0 + Wr[i * 3 + 1] * q1);
                Dp[i * 3 + 1] = Dm[i * 3 + 1] - Wm[i * 3 + 1] * Cp[2] - Wr[i * 3 + 1] * q2;
                Dp[i * 3 + 2] = Dm[i * 3 + 2] - Wm[i * 3 + 2] * Cp[3] - Wr[i * 3 + 2] * q3;
            }
            inv3x3(Dp, Di);
#pragma unroll
            for (int k = 0; k < 9; k++) Pm[k] = Wr[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) { Fa.Dinv[k][lane] = Di[k]; Fa.Wm[k][lane] = Wm[k]; Fa.Pm[k][lane] = Pm[k]; }
    // ---- 4. border: z = T^{-1} e_I, then the Schur complement ------------------------------------
    const double zf[3] = {0.0, 0.0, J.ps_I};   // border column e_I restricted to this node (Phi_s rows only)
    double u3[3];
    thomas_sweeps(m.Nx, ch, Wm, Pm, Di, zf, u3);
    Fa.z[0][lane] = u3[0]; Fa.z[1][lane] = u3[1]; Fa.z[2][lane] = u3[2];
    const double z0 = shfl_from(u3[2], 0), zN = shfl_from(u3[2], m.Nx - 1);
    if (lane == 0) {
        Fa.schur_inv = 1.0 / (ctrl.g_I - ctrl.g_ps0 * z0 - ctrl.g_psN * zN);
        Fa.g_ps0 = ctrl.g_ps0;
        Fa.g_psN = ctrl.g_psN;
    }
    __syncwarp();
}

// Solve J * d = g for one right-hand side held node-wise in registers (g in, d out, in place).
// gI: control-row right-hand side (uniform); returns dI (uniform).
__device__ __forceinline__ double warp_solve_impl(const ModelDesc& m, const LaneRole& ro, const WarpFactor& Fa,
                                                  bool alg_only, LaneVec& g, double gI, int lane) {
    const int el = ro.sec == 2 ? 1 : 0;
    // particle: s = Sinv * g_cs
    double s[NR];
    double p0 = 0.0;
    if (ro.elec && !alg_only) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < NR; c++) acc = fma(Fa.Sinv[r * NR + c][el], g.cs[c], acc);
            s[r] = acc;
        }
        p0 = s[NR - 1];
    } else {
#pragma unroll
        for (int r = 0; r < NR; r++) s[r] = 0.0;
    }
    const double inv_den = Fa.q[3][lane];
    const double q0 = ro.elec ? (g.j - Fa.jcs[lane] * p0) * inv_den : 0.0;
    double rf[3];
    rf[0] = alg_only ? 0.0 : g.ce - Fa.sj[0][lane] * q0;
    rf[1] = g.pe - Fa.sj[1][lane] * q0;
    rf[2] = (ro.elec ? g.ps : 0.0) - Fa.sj[2][lane] * q0;
    if (!ro.act) { rf[0] = rf[1] = rf[2] = 0.0; }
    double Di[9], Wm[9], Pm[9];
#pragma unroll
    for (int k = 0; k < 9; k++) { Di[k] = Fa.Dinv[k][lane]; Wm[k] = Fa.Wm[k][lane]; Pm[k] = Fa.Pm[k][lane]; }
    double u3[3];
    const LaneChain ch = make_chain(m.Nx, lane);
    thomas_sweeps(m.Nx, ch, Wm, Pm, Di, rf, u3);
    // border
    const double x0 = shfl_from(u3[2], 0), xN = shfl_from(u3[2], m.Nx - 1);
    const double dI = (gI - Fa.g_ps0 * x0 - Fa.g_psN * xN) * Fa.schur_inv;
    u3[0] -= Fa.z[0][lane] * dI; u3[1] -= Fa.z[1][lane] * dI; u3[2] -= Fa.z[2][lane] * dI;
    // back-substitute j and the particle
    const double dj = ro.elec ? q0 + Fa.q[0][lane] * u3[0] + Fa.q[1][lane] * u3[1] + Fa.q[2][lane] * u3[2] : 0.0;
    g.ce = alg_only ? 0.0 : u3[0];
    g.pe = u3[1];
    g.ps = ro.elec ? u3[2] : 0.0;
    g.j = dj;
#pragma unroll
    for (int r = 0; r < NR; r++) g.cs[r] = (ro.elec && !alg_only) ? s[r] - Fa.vb[r][el] * dj : 0.0;
    return dI;
}

// out-of-line copies for the operator-level Newton-init kernel
__device__ __noinline__ void warp_factor(const ModelDesc& m, const LaneRole& ro, const LaneJac& J,
                                         const CtrlRow& ctrl, double cj, bool alg_only, WarpFactor& Fa, int lane) {
    warp_factor_impl(m, ro, J, ctrl, cj, alg_only, Fa, lane);
}
__device__ __noinline__ double warp_solve(const ModelDesc& m, const LaneRole& ro, const WarpFactor& Fa,
                                          bool alg_only, LaneVec& g, double gI, int lane) {
    return warp_solve_impl(m, ro, Fa, alg_only, g, gI, lane);
}

}  // namespace plb
