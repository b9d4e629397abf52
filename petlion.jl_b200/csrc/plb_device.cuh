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
//
// This header is compiled twice (plb_variant_iso.cu: PLB_TH=0, namespace plb::iso; plb_variant_th.cu:
// PLB_TH=1, namespace plb::th).  With PLB_TH=1 the temperature is a state (temperature = true,
// residuals_T!: residuals.jl:299-489, build_heat_generation_rates!: auxiliary_states_and_coefficients.jl:
// 344-518): every lane also owns T of its node, lanes 0..Na-1 / Nx-Nz..Nx-1 own one current-collector
// node each, and the node blocks of the Newton matrix are 4x4 (c_e, Phi_e, Phi_s, T).
#pragma once
#include "plb_common.cuh"
#include "laws_generated.cuh"

#ifndef PLB_TH
#error "define PLB_TH (0/1) and PLB_NS before including the variant headers"
#endif

namespace plb {
namespace PLB_NS {

#ifndef PLB_SEI
#define PLB_SEI 0
#endif
#ifndef PLB_WIDE
#define PLB_WIDE 0
#endif
#ifndef PLB_MHC
#define PLB_MHC 0         // 1: the family also carries rxn_MHC (custom_functions.jl:233-298) next to rxn_BV.  Kept out of the default
#endif                    // builds: one more branch with an out-of-line call made the 168-register K1 spill (0.70 -> 0.53 of the roofline)
#ifndef PLB_DC
#define PLB_DC 0          // 1: the family also carries the concentration-rate inputs (METHOD_DC)
#endif
#if PLB_DC && (PLB_TH || PLB_SEI)
#error "the concentration-rate inputs are built for the isothermal families without aging"
#endif
constexpr bool TH = PLB_TH != 0;
constexpr bool SEI = PLB_SEI != 0;
// "Wide" families (grids with 33..64 x-nodes, e.g. N = (20,20,20)): one system is owned by a GROUP of two
// warps (64 lanes, lane x <-> node x as before).  Everything that is a warp shuffle / __syncwarp in the
// 32-lane families goes through the grp_* primitives below: in a wide family they exchange through a
// small shared-memory scratch area and a named barrier of the two warps, in the others they ARE the warp
// intrinsics (identical code generation).
constexpr bool WIDE = PLB_WIDE != 0;
constexpr int LW = WIDE ? 64 : 32;          // lanes per system
constexpr int XCH_K = 16;                   // doubles per lane of the exchange scratch (wide only)
constexpr size_t XCH_BYTES_PER_GROUP = WIDE ? sizeof(double) * XCH_K * LW : 0;
constexpr int NR = laws::NR;
// the j coupling of the particle rows in the eigen-basis, EVI * (b / cs_j): the surface column of EVI for the finite-difference
// scheme (b = cs_j e_surf), EVI * GJ for the spectral one
#if PLB_SPECTRAL
#define PLB_EVIB(i) laws::EVIG[i]
#else
#define PLB_EVIB(i) laws::EVI[i][NR - 1]
#endif

// ------------------------------------------------------------------------------------------------
// per-warp shared-memory constants derived from theta once per system
// ------------------------------------------------------------------------------------------------
enum SecField {
    SC_h, SC_inv_h, SC_inv_por, SC_pb, SC_Dlin, SC_src_ce, SC_hFa, SC_psf, SC_kap, SC_inv_Rp,
    SC_Rp_Ds, SC_k2, SC_cmax, SC_inv_cmax,
    // thermal: 1/(rho Cp), F a, sigma_eff, Ea_D/R, Ea_k/R   (SC_kap, SC_Rp_Ds, SC_k2 are then the values at T_ref)
    SC_irc, SC_Fa, SC_sig, SC_EaD, SC_Eak, SC_COUNT
};
enum GlobField { GC_T, GC_xcoef, GC_Kc, GC_I1C, GC_psI_p, GC_psI_n, GC_dUdT_on, GC_Tamb, GC_invL,
                 GC_xbc_a, GC_xbc_z,     // thermal: convective coefficient of the two outermost collector nodes
                
                 // aging = :SEI: R_SEI, 1/k_n_aging, Uref_s, i_0_jside/F, w, M_n/rho_n
                 GC_RSEI, GC_ikag, GC_Uref, GC_i0F, GC_w, GC_Mrho, GC_COUNT };

struct WarpConst {
    double sec[SC_COUNT][4];   // [field][section p,s,n,pad]
    double g[GC_COUNT + 1];
    double dinv[LW];           // 1 / (centre distance across face x|x+1)
    double beta[LW];           // harmonic-mean weight of face x|x+1
    double theta[TF_COUNT + 1];
#if PLB_TH
    // heat conduction (residuals.jl:299-446): coefficients of T[x-1]-T[x] and T[x+1]-T[x] in the T row of
    // node x, already divided by h*rho*Cp; same for this lane's current-collector node (xL, xR), its
    // convective boundary term xbc*(T_amb - T) and its Joule term xq*I^2
    double tL[LW], tR[LW], xL[LW], xR[LW], xq[LW];     // (xbc: two non-zero lanes, kept in g[GC_xbc_a / GC_xbc_z])
    double cinv[LW];           // 1 / (distance between the centres of nodes x-1 and x+1)
    double s5[3][8];           // h, lambda, rho*Cp of the five sections a,p,s,n,z
#endif
#if PLB_SEI
    double cSOH[LW];           // d(rhs_SOH)/d j_s of this lane's anode node (residuals.jl:278-297)
#endif
};

struct LaneRole {
    int x;        // node index == lane
    int sec;      // 0 p, 1 s, 2 n, 3 inactive
    int e;        // electrode index (p: x, n: x-Ns), -1 otherwise
    bool act, elec, first_e, last_e;   // first/last node of its electrode
    bool cha, chz;                     // owns a node of the positive / negative current collector (thermal)
    int ix;                            // offset of that node inside the T block, -1 otherwise
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
    r.cha = TH && lane < m.Na;
    r.chz = TH && lane >= m.Nx - m.Nz && lane < m.Nx;
    r.ix = r.cha ? lane : (r.chz ? m.Na + m.Nx + (lane - (m.Nx - m.Nz)) : -1);
    return r;
}

// one node's unknowns held in registers (T: temperature of the node, Tx: of this lane's current-
// collector node; both only exist in the thermal variant)
struct LaneVec {
    double ce, cs[NR], j, pe, ps;
    double T, Tx;
    double js, film, soh;      // aging = :SEI: side-reaction flux and film thickness (anode lanes), SOH (uniform)
};

// ---- group primitives ---------------------------------------------------------------------------
__device__ __forceinline__ int grp_lane() { return WIDE ? (int)(threadIdx.x & 63) : (int)(threadIdx.x & 31); }
__device__ __forceinline__ int grp_id() { return WIDE ? (int)(threadIdx.x >> 6) : (int)(threadIdx.x >> 5); }
__device__ __forceinline__ void grp_sync() {
#if PLB_WIDE
    asm volatile("bar.sync %0, 64;" ::"r"(1 + grp_id()) : "memory");
#else
    __syncwarp();
#endif
}
#if PLB_WIDE
// the exchange scratch of this thread's group: the first bytes of the CTA's dynamic shared memory
__device__ __forceinline__ double* grp_xch() {
    extern __shared__ __align__(16) unsigned char plb_dyn_smem[];
    return reinterpret_cast<double*>(plb_dyn_smem) + (size_t)grp_id() * XCH_K * LW;
}
#endif
__device__ __forceinline__ double shfl_from(double v, int src) {
#if PLB_WIDE
    double* x = grp_xch();
    x[grp_lane()] = v;
    grp_sync();
    const double r = x[src];
    grp_sync();
    return r;
#else
    return __shfl_sync(FULL, v, src);
#endif
}
// value of lane+1 / lane-1 (the end lanes get their own value back, like the warp intrinsics)
__device__ __forceinline__ double shfl_dn(double v) {
#if PLB_WIDE
    const int l = grp_lane();
    return shfl_from(v, l < LW - 1 ? l + 1 : l);
#else
    return __shfl_down_sync(FULL, v, 1);
#endif
}
__device__ __forceinline__ double shfl_up(double v) {
#if PLB_WIDE
    const int l = grp_lane();
    return shfl_from(v, l > 0 ? l - 1 : l);
#else
    return __shfl_up_sync(FULL, v, 1);
#endif
}
// value of lane+2 / lane-2 (the two end lanes get their own value back, like the warp intrinsics)
__device__ __forceinline__ double shfl_dn2(double v) {
#if PLB_WIDE
    const int l = grp_lane();
    return shfl_from(v, l < LW - 2 ? l + 2 : l);
#else
    return __shfl_down_sync(FULL, v, 2);
#endif
}
__device__ __forceinline__ double shfl_up2(double v) {
#if PLB_WIDE
    const int l = grp_lane();
    return shfl_from(v, l > 1 ? l - 2 : l);
#else
    return __shfl_up_sync(FULL, v, 2);
#endif
}
// K values from the same source lane in one exchange
template <int K>
__device__ __forceinline__ void shfl_from_n(const double* v, int src, double* out) {
#if PLB_WIDE
    static_assert(K <= XCH_K, "exchange scratch too small");
    double* x = grp_xch();
    const int l = grp_lane();
#pragma unroll
    for (int k = 0; k < K; k++) x[k * LW + l] = v[k];
    grp_sync();
#pragma unroll
    for (int k = 0; k < K; k++) out[k] = x[k * LW + src];
    grp_sync();
#else
#pragma unroll
    for (int k = 0; k < K; k++) out[k] = __shfl_sync(FULL, v[k], src);
#endif
}
__device__ __forceinline__ int grp_bcast_int(int v, int src) {
#if PLB_WIDE
    return (int)shfl_from((double)v, src);
#else
    return __shfl_sync(FULL, v, src);
#endif
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
#if PLB_WIDE
    {   // the two warps of the group: both add the partial sums in the same order
        double* x = grp_xch();
        if ((threadIdx.x & 31) == 0) x[(threadIdx.x >> 5) & 1] = v;
        grp_sync();
        v = x[0] + x[1];
        grp_sync();
    }
#endif
    return v;
}
__device__ __forceinline__ double grp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
#if PLB_WIDE
    {
        double* x = grp_xch();
        if ((threadIdx.x & 31) == 0) x[(threadIdx.x >> 5) & 1] = v;
        grp_sync();
        v = fmax(x[0], x[1]);
        grp_sync();
    }
#endif
    return v;
}
__device__ __forceinline__ double grp_min(double v) { return -grp_max(-v); }

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
    for (int f = lane; f < TF_COUNT; f += LW) {
        const int s = m.slot[f];
        C.theta[f] = s >= 0 ? theta_row[s] : 0.0;
    }
    grp_sync();
    const double* th = C.theta;
    // The per-section constants need ~20 reciprocals.  Computed by the three "section lanes" one after the other they
    // were a serial chain of ~450 instructions per system with 29 lanes idle (ncu source page of K1, round 1: 20 % of
    // the kernel's time); here every lane forms ONE of them (correctly rounded, __drcp_rn) and the section lanes only
    // multiply.  Reciprocal k (0..2: section p, s, n):  0-2 1/h   3-5 1/por   6-8 1/(rho Cp) [thermal]
    //   9,10 1/Rp (p, n)   11,12 1/sigma_eff   13,14 1/D_s   15,16 1/c_max   17 1/T0   18 1/(sum of the five lengths)
    double* const rc = C.beta;     // scratch: the face weights are written after the section constants have used it
    {
        const int k = lane, s = k < 9 ? k % 3 : (k - 9) & 1 ? 2 : 0;
        const double eps_f = s == 0 ? th[TF_eps_fp] : th[TF_eps_fn];
        const double eps_e = s == 0 ? th[TF_eps_p] : th[TF_eps_n];
        const double eps_act = 1.0 - (eps_f + eps_e);                                  // active_material
        double x = 1.0;
        if (k < 3) x = m.inv_n[s] * (s == 0 ? th[TF_l_p] : (s == 1 ? th[TF_l_s] : th[TF_l_n]));
        else if (k < 6) x = s == 1 ? th[TF_eps_s] : 1.0 - (eps_f + eps_act);            // build_eps!
        else if (k < 9) x = TH ? (s == 0 ? th[TF_rho_p] * th[TF_Cp_p] : (s == 1 ? th[TF_rho_s] * th[TF_Cp_s] : th[TF_rho_n] * th[TF_Cp_n])) : 1.0;
        else if (k < 11) x = s == 0 ? th[TF_Rp_p] : th[TF_Rp_n];
        else if (k < 13) x = (s == 0 ? th[TF_sigma_p] : th[TF_sigma_n]) * eps_act;      // build_sigma_eff
        else if (k < 15) x = s == 0 ? th[TF_D_sp] : th[TF_D_sn];
        else if (k < 17) x = s == 0 ? th[TF_c_max_p] : th[TF_c_max_n];
        else if (k == 17) x = th[TF_T0];
        else if (k == 18) x = TH ? th[TF_l_a] + th[TF_l_p] + th[TF_l_s] + th[TF_l_n] + th[TF_l_z] : 1.0;
        if (k < 19) rc[k] = __drcp_rn(x);
    }
    grp_sync();
    if (lane < 3) {
        const int s = lane, e = s == 2 ? 1 : 0;
        const double eps_f = s == 0 ? th[TF_eps_fp] : th[TF_eps_fn];
        const double eps_e = s == 0 ? th[TF_eps_p] : th[TF_eps_n];
        const double eps_act = 1.0 - (eps_f + eps_e);
        const double por = s == 1 ? th[TF_eps_s] : 1.0 - (eps_f + eps_act);
        const double brg = s == 0 ? th[TF_brugg_p] : (s == 1 ? th[TF_brugg_s] : th[TF_brugg_n]);
        const double l = s == 0 ? th[TF_l_p] : (s == 1 ? th[TF_l_s] : th[TF_l_n]);
        const double h = m.inv_n[s] * l;
        const double inv_Rp = rc[9 + e], inv_sig = rc[11 + e], inv_Ds0 = rc[13 + e];
        const double Rp = s == 0 ? th[TF_Rp_p] : th[TF_Rp_n];
        const double a = s == 1 ? 0.0 : 3 * eps_act * inv_Rp;                           // build_a!
        const double sig = (s == 0 ? th[TF_sigma_p] : th[TF_sigma_n]) * eps_act;
        const double Dl = s == 0 ? th[TF_D_p] : (s == 1 ? th[TF_D_s] : th[TF_D_n]);
        // por^brugg: the shipped parameter sets use brugg = 4 (LCO) and 1.5 (NMC); a general pow() is ~150
        // instructions that every evaluation of K1 would pay
        const double pb = brg == 4.0 ? (por * por) * (por * por) : (brg == 1.5 ? por * sqrt(por) : pow(por, brg));
        const double T = th[TF_T0];
        const bool Tref = (T == kTref);   // temperature_switch, custom_functions.jl:1
        double Ds = s == 0 ? th[TF_D_sp] : th[TF_D_sn];
        double inv_Ds = inv_Ds0;
        double k = s == 0 ? th[TF_k_p] : th[TF_k_n];
        if (!TH && !Tref && s != 1) {     // D_s_eff / rxn_rate Arrhenius, custom_functions.jl:16-57
            const double EaD = s == 0 ? th[TF_Ea_D_sp] : th[TF_Ea_D_sn];
            const double Eak = s == 0 ? th[TF_Ea_k_p] : th[TF_Ea_k_n];
            const double dre = rc[17] - 1.0 / kTref;
            const double arD = -EaD / kR * dre;
            Ds = Ds * exp(arD); inv_Ds = inv_Ds * exp(-arD);
            k = k * exp(-(Eak / kR) * dre);
        }
        const double cmax = s == 0 ? th[TF_c_max_p] : th[TF_c_max_n];
        C.sec[SC_h][s] = h;
        C.sec[SC_inv_h][s] = rc[s];
        C.sec[SC_inv_por][s] = rc[3 + s];
        C.sec[SC_pb][s] = pb;
        C.sec[SC_Dlin][s] = Dl * pb;
        C.sec[SC_src_ce][s] = (1 - th[TF_t_plus]) * a;
        C.sec[SC_hFa][s] = h * kF * a;
        C.sec[SC_psf][s] = s == 1 ? 0.0 : h * h * a * kF * inv_sig;
        C.sec[SC_kap][s] = s == 1 ? 0.0 : Ds * (inv_Rp * inv_Rp);
        C.sec[SC_inv_Rp][s] = s == 1 ? 0.0 : inv_Rp;
        C.sec[SC_Rp_Ds][s] = s == 1 ? 0.0 : Rp * inv_Ds;
        C.sec[SC_k2][s] = s == 1 ? 0.0 : 2.0 * k;
        C.sec[SC_cmax][s] = cmax;
        C.sec[SC_inv_cmax][s] = rc[15 + e];
        if (TH) {
            C.sec[SC_irc][s] = rc[6 + s];
            C.sec[SC_Fa][s] = kF * a;
            C.sec[SC_sig][s] = s == 1 ? 0.0 : sig;
            C.sec[SC_EaD][s] = s == 1 ? 0.0 : (s == 0 ? th[TF_Ea_D_sp] : th[TF_Ea_D_sn]) / kR;
            C.sec[SC_Eak][s] = s == 1 ? 0.0 : (s == 0 ? th[TF_Ea_k_p] : th[TF_Ea_k_n]) / kR;
        }
        if (s == 0) {
            const double I1C = calc_I1C(th);
            C.g[GC_T] = T;
            C.g[GC_xcoef] = (0.5 * kF / kR) * rc[17];
            C.g[GC_Kc] = (2 * kR / kF) * (1 - th[TF_t_plus]);
            C.g[GC_I1C] = I1C;
            C.g[GC_psI_p] = I1C * h * inv_sig;   // d res_Phi_s[first p] / dI   (residuals.jl:679)
            C.g[GC_dUdT_on] = (m.chem == CHEM_LCO && (TH || !Tref)) ? 1.0 : 0.0;
            if (SEI) {
                C.g[GC_RSEI] = th[TF_R_SEI]; C.g[GC_ikag] = 1.0 / th[TF_k_n_aging]; C.g[GC_Uref] = th[TF_Uref_s];
                C.g[GC_i0F] = th[TF_i_0_jside] / kF; C.g[GC_w] = th[TF_w]; C.g[GC_Mrho] = th[TF_M_n] / th[TF_rho_n];
            }
            if (TH) {
                C.g[GC_Tamb] = th[TF_T_amb];
                C.g[GC_invL] = rc[18];
            }
        }
        if (s == 2) C.g[GC_psI_n] = -calc_I1C(th) * h * inv_sig;   // d res_Phi_s[last n] / dI (residuals.jl:680)
    }
    grp_sync();
    {
        // face x | x+1 geometry: numerical_tools.jl:106-215
        const int x = lane;
        const int s0 = x < m.Np ? 0 : (x < m.Np + m.Ns ? 1 : 2);
        const int x1 = x + 1;
        const int s1 = x1 < m.Np ? 0 : (x1 < m.Np + m.Ns ? 1 : 2);
        const double h0 = C.sec[SC_h][s0], h1 = C.sec[SC_h][s1];
        double dinv, b;
        if (s0 == s1) { dinv = C.sec[SC_inv_h][s0]; b = 0.5; }
        else { dinv = __drcp_rn(h0 / 2 + h1 / 2); b = (h0 / 2) * dinv; }
        C.dinv[lane] = (x < m.Nx - 1) ? dinv : 0.0;
        C.beta[lane] = b;
    }
    grp_sync();
#if PLB_SEI
    {
        // residuals_SOH! (residuals.jl:278-297): rhs = F a_n/(3600 I1C) * trapz over the anode of j_s after a
        // quadratic extrapolation to both ends (extrapolate_section / extrap_x_0, external.jl:496-522): linear
        // in j_s, so every anode lane keeps its own weight
        // The geometry of that functional (the trapezoid weights with both extrapolated ends, for a section of unit
        // length) depends on N_n only: the host forms it once per model (ModelDesc::soh_geo); here it is scaled by l_n
        // and F a_n / (3600 I1C).
        const int k = lane - (m.Np + m.Ns);
        double wk = 0.0;
        if (k >= 0 && k < m.Nn) {
            const double eps_sn = 1.0 - (th[TF_eps_fn] + th[TF_eps_n]);
            wk = m.soh_geo[k] * th[TF_l_n] * (kF * (3 * eps_sn * C.sec[SC_inv_Rp][2]) / (3600 * C.g[GC_I1C]));
        }
        C.cSOH[lane] = wk;
    }
    grp_sync();
#endif
#if PLB_TH
    // ---- heat conduction coefficients, residuals.jl:299-446 (five sections a | p | s | n | z) -------
    if (lane < 5) {
        const int q = lane;
        const double l = q == 0 ? th[TF_l_a] : (q == 4 ? th[TF_l_z] : C.sec[SC_h][q - 1]);
        const int n = q == 0 ? m.Na : m.Nz;
        C.s5[0][q] = (q == 0 || q == 4) ? (1.0 / n) * l : l;
        C.s5[1][q] = q == 0 ? th[TF_lambda_a] : (q == 1 ? th[TF_lambda_p] : (q == 2 ? th[TF_lambda_s] : (q == 3 ? th[TF_lambda_n] : th[TF_lambda_z])));
        C.s5[2][q] = q == 0 ? th[TF_rho_a] * th[TF_Cp_a] : (q == 4 ? th[TF_rho_z] * th[TF_Cp_z] : 1.0 / C.sec[SC_irc][q - 1]);
    }
    grp_sync();
    {
        // conductance between the centres of two adjacent cells: lambda/h inside a section, harmonic-mean
        // lambda over the centre distance across an interface (residuals.jl:363-446).  Only 9 distinct conductances
        // (5 inside the sections a,p,s,n,z, 4 across their interfaces) and 5 scales 1/(h rho Cp) exist: lanes 0..13
        // form one each (a lane-local lambda with two divisions per call cost ~700 instructions per system)
        double* const G = C.xq;       // scratch [0..8] conductances, [9..13] scales; xq itself is written below
        if (lane < 14) {
            double v;
            if (lane < 5) v = C.s5[1][lane] / C.s5[0][lane];
            else if (lane < 9) {
                const int qa = lane - 5, qb = qa + 1;
                const double ha = C.s5[0][qa], hb = C.s5[0][qb], la = C.s5[1][qa], lb = C.s5[1][qb];
                const double be = (ha / 2) / (ha / 2 + hb / 2);
                const double lm = la * lb / (be * lb + (1.0 - be) * la);
                v = lm / (ha / 2 + hb / 2);
            } else v = 1.0 / (C.s5[0][lane - 9] * C.s5[2][lane - 9]);
            G[lane] = v;
        }
        grp_sync();
        double Gv[14];
#pragma unroll
        for (int k = 0; k < 14; k++) Gv[k] = G[k];
        grp_sync();                   // (every lane holds the table before xq is overwritten)
        auto cond = [&](int qa, int qb) -> double {
            // qa == qb: inside section qa; else the interface qa | qa+1
            double r = Gv[0];
#pragma unroll
            for (int k = 1; k < 9; k++) r = ((qa == qb ? qa : 5 + qa) == k) ? Gv[k] : r;
            return r;
        };
        auto scale = [&](int q) -> double {
            double r = Gv[9];
#pragma unroll
            for (int k = 1; k < 5; k++) r = (q == k) ? Gv[9 + k] : r;
            return r;
        };
        const int x = lane;
        auto sec5 = [&](int xx) -> int { return xx < 0 ? 0 : (xx < m.Np ? 1 : (xx < m.Np + m.Ns ? 2 : (xx < m.Nx ? 3 : 4))); };
        double tL = 0.0, tR = 0.0;
        if (x < m.Nx) {
            const int q = sec5(x);
            const double sc = scale(q);
            tL = cond(sec5(x - 1), q) * sc;
            tR = cond(q, sec5(x + 1)) * sc;
        }
        C.tL[lane] = tL; C.tR[lane] = tR;
        double xL = 0.0, xR = 0.0, xbc = 0.0, xq = 0.0;
        const double I1C = C.g[GC_I1C];
        if (lane < m.Na) {
            const int k = lane;
            const double sc = scale(0);
            xL = k > 0 ? cond(0, 0) * sc : 0.0;
            xR = (k < m.Na - 1 ? cond(0, 0) : cond(0, 1)) * sc;
            xbc = k == 0 ? th[TF_h_cell] * sc : 0.0;                       // T_BC_sx, residuals.jl:318
            xq = I1C * I1C / (th[TF_sigma_a] * C.s5[2][0]);                // residuals.jl:461
        } else if (lane >= m.Nx - m.Nz && lane < m.Nx) {
            const int k = lane - (m.Nx - m.Nz);
            const double sc = scale(4);
            xL = (k > 0 ? cond(4, 4) : cond(3, 4)) * sc;
            xR = k < m.Nz - 1 ? cond(4, 4) * sc : 0.0;
            xbc = k == m.Nz - 1 ? th[TF_h_cell] * sc : 0.0;               // T_BC_dx, residuals.jl:319
            xq = I1C * I1C / (th[TF_sigma_z] * C.s5[2][4]);                // residuals.jl:465
        }
        C.xL[lane] = xL; C.xR[lane] = xR; C.xq[lane] = xq;
        if (lane == 0) C.g[GC_xbc_a] = xbc;
        if (lane == m.Nx - 1) C.g[GC_xbc_z] = xbc;
        // central differences of thermal_derivatives (auxiliary_states_and_coefficients.jl:363-486):
        // (f[x+1]-f[x-1]) / (distance between the two centres), also across the section interfaces
        double ci = 0.0;
        if (x > 0 && x < m.Nx - 1) ci = 1.0 / (1.0 / C.dinv[x - 1] + 1.0 / C.dinv[x]);
        C.cinv[lane] = ci;
    }
    grp_sync();
#endif
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
#if PLB_SEI
    // aging = :SEI (anode lanes): the j row gains d/dj (through the film resistance) and d/dfilm; the j_s row;
    // the film row (d/dj_s; -cj on the diagonal); d(rhs_SOH)/dj_s.  The c_e, Phi_e, Phi_s rows see j + j_s,
    // so their d/dj_s equals their d/dj.
    double j_j, j_film, js_ps, js_pe, js_j, js_js, js_film, js_I, film_js, soh_js;
#endif
#if PLB_TH
    // temperature columns of the c_s, j and Phi_e rows
    double csT[NR], j_T, peTL, peTD, peTU;
    // T row of the node: conduction (T_TD without the -cj term), d/dj, d/d cs_surf, and the stencils of
    // thermal_derivatives over nodes x-2..x+2 (index 2 = own node) for c_e, Phi_e, Phi_s
    double T_TL, T_TD, T_TU, T_j, T_cs, T_ce[5], T_pe[5], T_ps[5];
    // this lane's current-collector node: tridiagonal (Tx_D without -cj) and d/dI
    double Tx_L, Tx_D, Tx_U, Tx_I;
#endif
#if PLB_TH && PLB_SEI
    // both: the side-reaction rate feels T (its Tafel exponent); the heat source F a j_total (T dU/dT + eta) feels j_s
    // like j, and j / film once more through the film-resistance term of eta
    double js_T, T_js, T_film;
#endif
};

// Control row (scalar_residual.jl:167-202): residual and its three possible Jacobian entries
struct CtrlRow {
    double res, g_ps0, g_psN, g_I;
    // dT control (thermal variant): coefficients of this lane's Y'[T] / Y'[collector T] in the row
    // (-temperature_weighting weights); zero for the other methods
    double gTn, gTx;
    // eta_p control: +g_eta on Phi_s and -g_eta on Phi_e of the first anode node; zero for the other methods
    double g_eta;
#if PLB_DC
    // concentration-rate control (METHOD_DC / METHOD_DC_ALG): the row holds one state of lane dc_tgt
    int dc_on, dc_tgt, dc_comp, dc_alg;
#endif
};

// rxn_MHC (custom_functions.jl:233-298, the alpha == 0.5 branch -- the only live one): value and partial derivatives
// with respect to (eta, c_e, c_s_star) at fixed k and T; fT = F/(R T).  Every sqrt of the reference function is
// sqrt_ReLU, log_ReLU(x; 1e-4) = log(max(1e-4, x)).
#if PLB_MHC
struct MhcRate { double j, d_eta, d_ce, d_cs; };
__device__ __noinline__ void mhc_rate(double cs, double ce, double eta, double k_i, double fT, double lam, double cmax,
                                      double ce0, bool with_jac, MhcRate& o) {
    const double ratio = (ce / ce0) / (cs / cmax);
    const bool live = ratio > 1e-4;
    const double eta_f = eta * fT + (live ? log(ratio) : log(1e-4));
    const double sl = sqrt(fmax(lam, 0.0));
    const double a = 1.0 + sl;
    const double k0 = k_i / ((1.0 - erf((lam - sqrt(a)) / (2.0 * sl))) / 2.0);
    const double r = sqrt(fmax(a + eta_f * eta_f, 0.0));
    const double u = (lam - r) / (2.0 * sl);
    const double g = 1.0 - erf(u);
    const double em = exp(-eta_f);
    const double sp = 1.0 / (1.0 + em), sm = 1.0 / (1.0 + exp(eta_f));
    const double B = sp * ce0 * cs - sm * ce * cmax;
    const double sa = (1.0 - cs / cmax) / ce0;
    const double Sq = sa > 0.0 ? sqrt(sa) : 0.0;
    o.j = k0 * g * B * Sq;
    if (with_jac) {
        // d g / d eta_f = (2/sqrt(pi)) exp(-u^2) (eta_f / r) / (2 sl)
        const double dg = r > 0.0 ? 1.1283791670955126 * exp(-u * u) * eta_f / (r * 2.0 * sl) : 0.0;
        const double dB = (ce0 * cs + ce * cmax) * sp * sm;
        const double dj_f = k0 * Sq * (dg * B + g * dB);                 // d j / d eta_f
        o.d_eta = dj_f * fT;
        o.d_ce = (live ? dj_f / ce : 0.0) - k0 * g * Sq * cmax * sm;
        o.d_cs = (live ? -dj_f / cs : 0.0) + k0 * g * Sq * ce0 * sp + (Sq > 0.0 ? -k0 * g * B / (2.0 * Sq * cmax * ce0) : 0.0);
    }
}
#endif

template <int CHEM, bool WITH_JAC>
__device__ __forceinline__ void lane_eval(const ModelDesc& m, const WarpConst& C, const LaneRole& ro,
                                          const LaneVec& y, const LaneVec& yp, double Iapp,
                                          int method, double value, LaneVec& res, CtrlRow& ctrl,
                                          LaneJac& J) {
#if PLB_DC
    const int dc_tgt = (method >> 8) & 0xff, dc_comp = (method >> 16) & 1;
    method &= 0xff;
#endif
    const int s = ro.sec < 3 ? ro.sec : 1;
    const double T = TH ? (ro.act ? y.T : kTref) : C.g[GC_T];
    // ---- node-local electrolyte properties: build_K_eff!/build_D_eff! (:302-328) -----------------
    double K = 0.0, dK = 0.0, D = 0.0, dD = 0.0, dKT = 0.0;
    const double ce = ro.act ? y.ce : 1000.0;
    {
        const double pb = C.sec[SC_pb][s];
        if (CHEM == CHEM_LGM) {
            // K_eff_LGM50, D_eff_LGM50 (params.jl:648-672): functions of c_e only
            laws::K_eff_LGM50(ce, sqrt(fmax(ce, 0.0) * 1e-3), K, dK);
            K *= pb; dK *= pb;
            laws::D_eff_LGM50(ce, D, dD);
            const double De = C.theta[TF_D_e] * pb;
            D *= De; dD *= De;
        } else {
            if (TH) { laws::K_eff_T(ce, T, K, dK, dKT); dKT *= pb; }
            else laws::K_eff(ce, T, K, dK);
            K *= pb; dK *= pb;
            if (CHEM == CHEM_LCO) { D = C.sec[SC_Dlin][s]; dD = 0.0; }   // D_eff_linear
            else { laws::D_eff_nl(ce, T, D, dD); D *= pb; dD *= pb; }
        }
    }
    // ---- face x|x+1 (owned by lane x): harmonic means and fluxes ---------------------------------
    const double ceR = shfl_dn(ce), peR = shfl_dn(y.pe), KR = shfl_dn(K), dKR = shfl_dn(dK),
                 DR = shfl_dn(D), dDR = shfl_dn(dD);
    const bool has_face = ro.x < m.Nx - 1;
    const double b = C.beta[ro.x], dinv = C.dinv[ro.x];
    double Nf = 0.0, Q = 0.0;
    double dN_cL = 0.0, dN_cR = 0.0, dQ_cL = 0.0, dQ_cR = 0.0, wK = 0.0;
    double dQ_TL = 0.0, dQ_TR = 0.0;
    double TRn = T, dKTR = 0.0;
    if (TH) { TRn = shfl_dn(T); dKTR = shfl_dn(dKT); }
    if (has_face) {
        // harmonic means (numerical_tools.jl:106-191); one reciprocal per mean, shared with its derivatives
        const double denK = b * KR + (1.0 - b) * K;
        const double rK = __drcp_rn(denK);
        const double Khat = K * KR * rK;                            // interpolate_electrolyte_grid
        const double denD = b * DR + (1.0 - b) * D;
        const double rD = __drcp_rn(denD);
        const double Dhat = D * DR * rD;
        const double denc = b * ceR + (1.0 - b) * ce;
        const double rc = __drcp_rn(denc);
        const double cbar = ce * ceR * rc;                          // interpolate_electrolyte_concentration
        const double denT = b * TRn + (1.0 - b) * T;
        const double Tbar = TH ? T * TRn * __drcp_rn(denT) : T;     // interpolate_temperature (uniform T: the mean is T)
        const double dc = (ceR - ce) * dinv;                        // ..._concetration_fluxes
        const double icb = __drcp_rn(cbar);
        const double G = Khat * Tbar * dc * icb;                    // prod_tot, residuals.jl:631-635
        wK = Khat * dinv;
        Q = wK * (peR - y.pe) - C.g[GC_Kc] * G;
        Nf = Dhat * dinv * (ceR - ce);
        if (WITH_JAC) {
            const double iK2 = rK * rK, iD2 = rD * rD, ic2 = rc * rc;
            const double dKh_cL = b * KR * KR * iK2 * dK, dKh_cR = (1.0 - b) * K * K * iK2 * dKR;
            const double dDh_cL = b * DR * DR * iD2 * dD, dDh_cR = (1.0 - b) * D * D * iD2 * dDR;
            const double dcb_cL = b * ceR * ceR * ic2, dcb_cR = (1.0 - b) * ce * ce * ic2;
            const double dG_cL = Tbar * icb * (dKh_cL * dc - Khat * dinv - Khat * dc * dcb_cL * icb);
            const double dG_cR = Tbar * icb * (dKh_cR * dc + Khat * dinv - Khat * dc * dcb_cR * icb);
            dQ_cL = dKh_cL * dinv * (peR - y.pe) - C.g[GC_Kc] * dG_cL;
            dQ_cR = dKh_cR * dinv * (peR - y.pe) - C.g[GC_Kc] * dG_cR;
            dN_cL = -Dhat * dinv + dDh_cL * dc;
            dN_cR = Dhat * dinv + dDh_cR * dc;
            if (TH) {
                const double iT2 = 1.0 / (denT * denT);
                const double dKh_TL = b * KR * KR * iK2 * dKT, dKh_TR = (1.0 - b) * K * K * iK2 * dKTR;
                const double dTb_TL = b * TRn * TRn * iT2, dTb_TR = (1.0 - b) * T * T * iT2;
                const double dcc = dc * icb;
                dQ_TL = dKh_TL * dinv * (peR - y.pe) - C.g[GC_Kc] * dcc * (dKh_TL * Tbar + Khat * dTb_TL);
                dQ_TR = dKh_TR * dinv * (peR - y.pe) - C.g[GC_Kc] * dcc * (dKh_TR * Tbar + Khat * dTb_TR);
            }
        }
    }
    // left face x-1|x comes from lane x-1
    double NfL = shfl_up(Nf), QL = shfl_up(Q), wKL = shfl_up(wK);
    double dNL_cL = 0.0, dNL_cR = 0.0, dQL_cL = 0.0, dQL_cR = 0.0;
    double dQL_TL = 0.0, dQL_TR = 0.0;
    if (WITH_JAC) { dNL_cL = shfl_up(dN_cL); dNL_cR = shfl_up(dN_cR); dQL_cL = shfl_up(dQ_cL); dQL_cR = shfl_up(dQ_cR); }
    if (WITH_JAC && TH) { dQL_TL = shfl_up(dQ_TL); dQL_TR = shfl_up(dQ_TR); }
    if (ro.x == 0) { NfL = 0.0; QL = 0.0; wKL = 0.0; dNL_cL = dNL_cR = dQL_cL = dQL_cR = 0.0; dQL_TL = dQL_TR = 0.0; }

    // ---- electrode-node quantities ------------------------------------------------------------------
    double jtot = 0.0, jcalc = 0.0;
    double dj_cs = 0.0, dj_ce = 0.0, dj_eta = 0.0, dj_T = 0.0;
    // thermal: Arrhenius factors of D_s_eff / rxn_rate at the node temperature (custom_functions.jl:16-57),
    // surface OCV data for the heat sources
    double kapx = C.sec[SC_kap][s], k2x = C.sec[SC_k2][s], xco = C.g[GC_xcoef], RpDs = C.sec[SC_Rp_Ds][s];
    double dkapT = 0.0, eta_h = 0.0, dUdT_h = 0.0, ddUdT_h = 0.0, dUtot_h = 0.0;
    if (TH && ro.elec) {
        const double iT = 1.0 / T;
        const double dre = iT - 1.0 / kTref;
        const double aD = exp(-C.sec[SC_EaD][s] * dre), ak = exp(-C.sec[SC_Eak][s] * dre);
        kapx *= aD; RpDs /= aD; k2x *= ak;
        dkapT = kapx * C.sec[SC_EaD][s] * iT * iT;
        xco = 0.5 * kF / (kR * T);
    }
    const bool sei_n = SEI && ro.sec == 2;
    double Rfilm = 0.0;
    if (ro.elec) {
        jtot = sei_n ? y.j + y.js : y.j;                                     // build_j_total!
        if (sei_n) Rfilm = C.g[GC_RSEI] + y.film * C.g[GC_ikag];
        const double cs_s = y.cs[NR - 1];                                    // build_c_s_star!
        const double th = cs_s * C.sec[SC_inv_cmax][s];
        double U, dU, dUdT = 0.0, ddUdT = 0.0;
        if (CHEM == CHEM_LCO) {
            // sqrt_ReLU branches (custom_functions.jl:143, 210): physical range th > 1e-4
            const double sq = sqrt(fmax(th, 1e-4));
            if (TH || C.g[GC_dUdT_on] != 0.0) {
                if (ro.sec == 0) laws::OCV_LCO(th, U, dU, dUdT, ddUdT);
                else laws::OCV_LiC6(th, sq, U, dU, dUdT, ddUdT);
                U += dUdT * (T - kTref); dU += ddUdT * (T - kTref);
            } else {
                // isothermal run at T == T_ref: the entropic term is switched off (temperature_switch,
                // custom_functions.jl:1), skip its two rational polynomials
                if (ro.sec == 0) laws::OCV_LCO_U(th, U, dU);
                else laws::OCV_LiC6_U(th, sq, U, dU);
            }
        } else if (CHEM == CHEM_LGM) {
            if (ro.sec == 0) laws::OCV_NMC811(th, U, dU);           // (no entropic term: dU/dT = 0, params.jl:564, 637)
            else laws::OCV_LiC6_LGM50(th, U, dU);
        } else {
            if (ro.sec == 0) laws::OCV_NMC(th, U, dU);
            else laws::OCV_LiC6_NMC(th, U, dU);
        }
        const double eta = y.ps - y.pe - U - (sei_n ? kF * y.j * Rfilm : 0.0);   // build_eta! (:272-300)
        const double cmax = C.sec[SC_cmax][s];
#if PLB_MHC
        if (m.rxn_mhc != 0 && ((m.rxn_mhc >> (ro.sec == 2 ? 1 : 0)) & 1)) {
            // rxn_MHC, custom_functions.jl:233-298 (out of line: a model option, kept off the registers of rxn_BV)
            MhcRate o;
            mhc_rate(cs_s, ce, eta, 0.5 * k2x, 2.0 * xco, C.theta[ro.sec == 2 ? TF_lambda_MHC_n : TF_lambda_MHC_p], cmax,
                     C.theta[TF_c_e0], WITH_JAC, o);
            jcalc = o.j;
            if (TH) { eta_h = eta; dUdT_h = dUdT; ddUdT_h = ddUdT; dUtot_h = dU; }
            if (WITH_JAC) {
                dj_eta = o.d_eta;
                dj_ce = o.d_ce;
                dj_cs = o.d_cs - dj_eta * dU * C.sec[SC_inv_cmax][s];
#if PLB_SEI
                J.j_j = -1.0 - (sei_n ? dj_eta * kF * Rfilm : 0.0);
                J.j_film = sei_n ? -dj_eta * kF * y.j * C.g[GC_ikag] : 0.0;
#endif
                if (TH) {
                    const double iT = 1.0 / T;       // k(T), U(T), and the 1/T of eta_hat
                    dj_T = jcalc * C.sec[SC_Eak][s] * iT * iT - dj_eta * dUdT - dj_eta * eta * iT;
                }
            }
        } else
#endif
        {
            // rxn_BV, custom_functions.jl:212-231
            const double arg = ce * cs_s * (cmax - cs_s);
            const double sq = arg > 0.0 ? sqrt(arg) : 0.0;                       // sqrt_ReLU
            const double xx = xco * eta;
            const double em = expm1(xx);
            const double rem = __drcp_rn(em + 1.0);                              // exp(-xx)
            const double sh = 0.5 * (em + em * rem);                             // sinh(xx)
            const double k2 = k2x;
            jcalc = k2 * sq * sh;
            if (TH) { eta_h = eta; dUdT_h = dUdT; ddUdT_h = ddUdT; dUtot_h = dU; }
            if (WITH_JAC) {
                const double ch = sh + rem;                                      // cosh = sinh + exp(-x)
                const double isq = arg > 0.0 ? 0.5 / sq : 0.0;
                dj_eta = k2 * sq * ch * xco;
                dj_ce = k2 * sh * isq * cs_s * (cmax - cs_s);
                dj_cs = k2 * sh * isq * ce * (cmax - 2.0 * cs_s) - dj_eta * dU * C.sec[SC_inv_cmax][s];
#if PLB_SEI
                J.j_j = -1.0 - (sei_n ? dj_eta * kF * Rfilm : 0.0);
                J.j_film = sei_n ? -dj_eta * kF * y.j * C.g[GC_ikag] : 0.0;
#endif
                if (TH) {
                    // d/dT: k(T), eta(T) through U = U0 + dUdT (T - Tref), and the 1/T in the sinh argument
                    const double iT = 1.0 / T;
                    dj_T = jcalc * C.sec[SC_Eak][s] * iT * iT + k2 * sq * ch * (-xco * dUdT - xx * iT);
                }
            }
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
        const double kap = kapx;
        const double d1bc = -y.j * RpDs;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < NR; c++)
                if (laws::mc_mask(r) & (1u << c)) acc = fma(laws::MC[r][c], y.cs[c], acc);
#if PLB_TH
            if (WITH_JAC) J.csT[r] = dkapT * acc;     // kap*BJ*d1bc = -BJ*j/Rp does not depend on T
#endif
#if PLB_SPECTRAL
            acc = fma(laws::GJ[r], d1bc, acc);        // Chebyshev collocation: the surface flux reaches every radial row
#else
            if (r == NR - 1) acc = fma(laws::BJ, d1bc, acc);
#endif
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
#if PLB_TH
        if (WITH_JAC) {
#pragma unroll
            for (int r = 0; r < NR; r++) J.csT[r] = 0.0;
        }
#endif
    }
#if PLB_SEI
    // ---- aging = :SEI: residuals_j_s! (residuals.jl:519-552), residuals_film! (:260-276), residuals_SOH! (:278-297)
    {
        double jsc = 0.0, Dn = 0.0, eta_s = 0.0, xs = 0.0;
        const bool charging = Iapp * C.g[GC_I1C] > 0.0;                   // I_density > 0, :546
        if (sei_n && charging) {
            eta_s = y.ps - y.pe - C.g[GC_Uref] - kF * jtot * Rfilm;
            xs = 0.5 * kF / (kR * T);
            const double Ir = Iapp * C.g[GC_I1C] / C.g[GC_I1C];
            jsc = -fabs(C.g[GC_i0F] * pow(Ir, C.g[GC_w]) * (-exp(-xs * eta_s)));
            Dn = -xs * jsc;                                               // d j_s_calc / d eta_s
        }
        res.js = sei_n ? y.js - jsc : 0.0;
        res.film = sei_n ? -y.js * C.g[GC_Mrho] - yp.film : 0.0;
        res.soh = warp_sum(sei_n ? C.cSOH[ro.x] * y.js : 0.0) - yp.soh;
        if (WITH_JAC) {
            J.js_ps = -Dn; J.js_pe = Dn;
            J.js_j = Dn * kF * Rfilm;
            J.js_js = sei_n ? 1.0 + Dn * kF * Rfilm : 1.0;
            J.js_film = Dn * kF * jtot * C.g[GC_ikag];
            J.js_I = (sei_n && charging) ? -C.g[GC_w] * jsc / Iapp : 0.0;
            J.film_js = sei_n ? -C.g[GC_Mrho] : 0.0;
            J.soh_js = sei_n ? C.cSOH[ro.x] : 0.0;
            if (!ro.elec) { J.j_j = -1.0; J.j_film = 0.0; }
#if PLB_TH
            J.js_T = -jsc * eta_s * xs / T;                               // -(d j_s_calc / dT): the 1/T of the exponent
#endif
        }
    }
#endif
#if PLB_TH
    // ---- residuals_T!, residuals.jl:299-489; heat sources auxiliary_states_and_coefficients.jl:344-518 --
    {
        // thermal_derivatives: one-sided 3-point differences at the outer ends of the cell (c_e, Phi_e) and
        // of each electrode (Phi_s), central differences elsewhere.  w*[k] multiplies the value at x-2+k.
        const bool x0 = ro.x == 0, xN = ro.x == m.Nx - 1;
        const double h2 = 0.5 * C.sec[SC_inv_h][s], ci = C.cinv[ro.x];
        double we[5], ws[5];
        we[0] = xN ? h2 : 0.0;
        we[1] = xN ? -4.0 * h2 : (x0 ? 0.0 : -ci);
        we[2] = x0 ? -3.0 * h2 : (xN ? 3.0 * h2 : 0.0);
        we[3] = x0 ? 4.0 * h2 : (xN ? 0.0 : ci);
        we[4] = x0 ? -h2 : 0.0;
        ws[0] = (ro.elec && ro.last_e) ? h2 : 0.0;
        ws[1] = !ro.elec ? 0.0 : (ro.last_e ? -4.0 * h2 : (ro.first_e ? 0.0 : -h2));
        ws[2] = !ro.elec ? 0.0 : (ro.first_e ? -3.0 * h2 : (ro.last_e ? 3.0 * h2 : 0.0));
        ws[3] = !ro.elec ? 0.0 : (ro.first_e ? 4.0 * h2 : (ro.last_e ? 0.0 : h2));
        ws[4] = (ro.elec && ro.first_e) ? -h2 : 0.0;
        if (!ro.act) {
#pragma unroll
            for (int k = 0; k < 5; k++) { we[k] = 0.0; ws[k] = 0.0; }
        }
        const double ceL1 = shfl_up(ce), ceL2 = shfl_up2(ce), ceR2 = shfl_dn2(ce);
        const double peL1 = shfl_up(y.pe), peL2 = shfl_up2(y.pe), peR2 = shfl_dn2(y.pe);
        const double psL2 = shfl_up2(y.ps), psR2 = shfl_dn2(y.ps);
        const double dPe = we[0] * peL2 + we[1] * peL1 + we[2] * y.pe + we[3] * peR + we[4] * peR2;
        const double dCe = we[0] * ceL2 + we[1] * ceL1 + we[2] * ce + we[3] * ceR + we[4] * ceR2;
        const double dPs = ws[0] * psL2 + ws[1] * psL + ws[2] * y.ps + ws[3] * psR + ws[4] * psR2;
        const double Kc = C.g[GC_Kc], irc = C.sec[SC_irc][s], Fa = C.sec[SC_Fa][s], sig = C.sec[SC_sig][s];
        const double ice = 1.0 / ce;
        const double Qohm = K * dPe * dPe + Kc * K * T * (dCe * ice) * dPe + sig * dPs * dPs;
        const double Qrr = ro.elec ? Fa * jtot * (T * dUdT_h + eta_h) : 0.0;     // Q_rev + Q_rxn
        // neighbours of the temperature stencil: the two end nodes talk to the current collectors
        double TL = shfl_up(T), TR = TRn;
        const double Ta_last = shfl_from(y.Tx, m.Na - 1), Tz_first = shfl_from(y.Tx, m.Nx - m.Nz);
        if (x0) TL = Ta_last;
        if (xN) TR = Tz_first;
        const double tL = C.tL[ro.x], tR = C.tR[ro.x];
        res.T = ro.act ? tL * (TL - T) + tR * (TR - T) + irc * (Qohm + Qrr) - yp.T : 0.0;
        // current-collector node of this lane
        const double TxU = shfl_up(y.Tx), TxD = shfl_dn(y.Tx);
        const double T_first = shfl_from(T, 0), T_last = shfl_from(T, m.Nx - 1);
        double TxL = TxU, TxR = TxD;
        if (ro.cha && ro.x == m.Na - 1) TxR = T_first;
        if (ro.chz && ro.x == m.Nx - m.Nz) TxL = T_last;
        const double xL = C.xL[ro.x], xR = C.xR[ro.x], xq = C.xq[ro.x];
        const double xbc = (ro.cha && ro.x == 0) ? C.g[GC_xbc_a] : ((ro.chz && ro.x == m.Nx - 1) ? C.g[GC_xbc_z] : 0.0);
        res.Tx = (ro.cha || ro.chz)
                     ? xL * (TxL - y.Tx) + xR * (TxR - y.Tx) + xq * Iapp * Iapp + xbc * (C.g[GC_Tamb] - y.Tx) - yp.Tx
                     : 0.0;
        if (WITH_JAC) {
            J.T_TL = tL; J.T_TU = tR;
            J.T_TD = -(tL + tR) + irc * (dKT * dPe * dPe + Kc * dPe * (dCe * ice) * (dKT * T + K));
            J.T_j = ro.elec ? irc * Fa * (T * dUdT_h + eta_h) : 0.0;
#if PLB_SEI
            J.T_js = sei_n ? J.T_j : 0.0;
            J.T_film = sei_n ? -irc * Fa * jtot * kF * y.j * C.g[GC_ikag] : 0.0;      // through eta's film-resistance term
            if (sei_n) J.T_j -= irc * Fa * jtot * kF * Rfilm;
#endif
            J.T_cs = ro.elec ? irc * Fa * jtot * (T * ddUdT_h - dUtot_h) * C.sec[SC_inv_cmax][s] : 0.0;
            const double gpe = irc * (2.0 * K * dPe + Kc * K * T * dCe * ice);
            const double gce = irc * Kc * K * T * dPe * ice;
            const double gps = irc * 2.0 * sig * dPs;
#pragma unroll
            for (int k = 0; k < 5; k++) { J.T_pe[k] = gpe * we[k]; J.T_ce[k] = gce * we[k]; J.T_ps[k] = gps * ws[k]; }
            const double src_j = ro.elec ? irc * Fa * jtot : 0.0;      // d(Q_rxn)/d eta
            J.T_pe[2] -= src_j;
            J.T_ps[2] += src_j;
            J.T_ce[2] += irc * (dK * dPe * dPe + Kc * T * dPe * dCe * (dK * ice - K * ice * ice));
            if (!ro.act) { J.T_TL = 0.0; J.T_TU = 0.0; J.T_TD = 0.0; }
            J.Tx_L = xL; J.Tx_U = xR; J.Tx_D = -(xL + xR) - xbc; J.Tx_I = 2.0 * xq * Iapp;
        }
    }
#endif
    // ---- control row: scalar_residual!, scalar_residual.jl:167; calc_V/P :86-87 ----------------------
    {
        const double ps0 = shfl_from(y.ps, 0), psN = shfl_from(y.ps, m.Nx - 1);
        const double V = ps0 - psN;
        ctrl.gTn = 0.0; ctrl.gTx = 0.0; ctrl.g_eta = 0.0;
#if PLB_DC
        ctrl.dc_on = 0; ctrl.dc_tgt = dc_tgt; ctrl.dc_comp = dc_comp; ctrl.dc_alg = 0;
        if (method == METHOD_DC || method == METHOD_DC_ALG) {
            // run_residual of state_deriv_func(ind): val - Y'[ind] (scalar_residual.jl:172, auxiliary_states_and_coefficients.jl:682);
            // inside newtons_method! Y'[ind] is replaced by the right-hand side of that state's own row (:347-363)
            const bool alg = method == METHOD_DC_ALG;
            const double mine = dc_comp ? (alg ? res.ce + yp.ce : yp.ce) : (alg ? res.cs[NR - 1] + yp.cs[NR - 1] : yp.cs[NR - 1]);
            ctrl.res = value - shfl_from(mine, dc_tgt);
            ctrl.g_ps0 = 0.0; ctrl.g_psN = 0.0; ctrl.g_I = 0.0;
            ctrl.dc_on = 1; ctrl.dc_alg = alg ? 1 : 0;
        } else
#endif
        if (method == METHOD_ETA) {
            const int ln = m.Np + m.Ns;                                  // first anode node
            ctrl.res = (shfl_from(y.ps, ln) - shfl_from(y.pe, ln)) - value;   // calc_eta_plating, scalar_residual.jl:92
            ctrl.g_ps0 = 0.0; ctrl.g_psN = 0.0; ctrl.g_I = 0.0; ctrl.g_eta = 1.0;
        } else if (method == METHOD_I) { ctrl.res = Iapp - value; ctrl.g_ps0 = 0.0; ctrl.g_psN = 0.0; ctrl.g_I = 1.0; }
#if PLB_TH
        else if (method == METHOD_DT || method == METHOD_DT_ALG) {
            // run_residual of the constant-temperature mode: val - temperature_weighting(Y'[T])
            // (scalar_residual.jl:172, auxiliary_states_and_coefficients.jl:649-679)
            const double invL = C.g[GC_invL];
            const double wn = ro.act ? C.s5[0][1 + s] * invL : 0.0;
            const double wc = ro.cha ? C.s5[0][0] * invL : (ro.chz ? C.s5[0][4] * invL : 0.0);
            const bool alg = method == METHOD_DT_ALG;
            const double a = alg ? wn * (res.T + yp.T) + wc * (res.Tx + yp.Tx) : wn * yp.T + wc * yp.Tx;
            ctrl.res = value - warp_sum(a);
            ctrl.g_ps0 = 0.0; ctrl.g_psN = 0.0; ctrl.gTn = -wn; ctrl.gTx = -wc;
            // inside newtons_method! the row depends on I through the Joule heating of the collectors
            ctrl.g_I = alg ? -warp_sum(wc * 2.0 * C.xq[ro.x] * Iapp) : 0.0;
        }
#endif
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
        J.kap = kapx;
        J.cs_j = ro.elec ? -laws::BJ * C.sec[SC_inv_Rp][s] : 0.0;     // (spectral: BJ = 1, row r carries cs_j * GJ[r])
#if PLB_TH
        J.j_T = dj_T;
        if (last) { J.peTL = 0.0; J.peTD = 0.0; J.peTU = 0.0; }
        else { J.peTL = dQL_TL; J.peTD = dQL_TR - dQ_TL; J.peTU = -dQ_TR; }
#endif
    }
}

// out-of-line copy for the fused integrator (keeps its code inside the instruction cache)
template <int CHEM, bool WITH_JAC>
__device__ __noinline__ void lane_eval_ni(const ModelDesc& m, const WarpConst& C, const LaneRole& ro,
                                          const LaneVec& y, const LaneVec& yp, double Iapp, int method,
                                          double value, LaneVec& res, CtrlRow& ctrl, LaneJac& J) {
    lane_eval<CHEM, WITH_JAC>(m, C, ro, y, yp, Iapp, method, value, res, ctrl, J);
}

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

#if !PLB_TH
// ------------------------------------------------------------------------------------------------
// structured Newton-matrix factorisation / solve
//   1. particle block (kap*MC - cj I) is identical for every particle of an electrode -> explicit
//      10x10 inverse per electrode, formed from the stencil's constant eigen-decomposition: EV diag(1/(kap*EL-cj)) EVI
//   2. eliminate c_s (through its surface value) and j node-locally
//   3. 3x3-block tridiagonal system in (c_e, Phi_e, Phi_s) over the nodes: block Thomas along lanes
//   4. the applied-current unknown I is a border (Schur complement), which also covers the
//      zero-diagonal control row of voltage/power control (scalar_residual.jl:184-197)
// ------------------------------------------------------------------------------------------------
// The factored 3x3 blocks (Dinv, W, P: 27 doubles per lane, 6.9 KB per system) live in the system's slot of the GLOBAL workspace
// (L2-resident), like the thermal families' 4x4 blocks: written once per factorisation, read once at the head of a solve (27
// independent coalesced loads per lane).  The shared memory they occupied holds phi_4 / phi_5 again (PLB_NGLOBAL = 0), which the
// BDF vector passes touch far more often.
#ifndef PLB_BLOCKS_GLOBAL
#define PLB_BLOCKS_GLOBAL 1
#endif
// aging = :SEI: the 3x3 local inverse Mi and the six couplings cpl of the (j, j_s, film) elimination live there as well (15 doubles
// per lane, read once at the head of a solve): the room for a fourth system per SM on the two-warp grid (cfg5)
#ifndef PLB_SEI_LOCAL_GLOBAL
#define PLB_SEI_LOCAL_GLOBAL (PLB_SEI && PLB_WIDE && PLB_BLOCKS_GLOBAL)     // (32-node SEI: seven systems fit either way; on chip: 368 k vs 362 k)
#endif
constexpr int FA_GLOBAL = PLB_BLOCKS_GLOBAL ? (27 + (PLB_SEI_LOCAL_GLOBAL ? 15 : 0)) * LW : 0;      // doubles per system slot of the global workspace
#if PLB_SEI_LOCAL_GLOBAL
#define FA_MI(k) Fa.blk[(27 + (k)) * LW + lane]
#define FA_CPL(k) Fa.blk[(36 + (k)) * LW + lane]
#else
#define FA_MI(k) Fa.Mi[k][lane]
#define FA_CPL(k) Fa.cpl[k][lane]
#endif
struct WarpFactor {
    double Sinv[NR * NR][2];   // [r*NR+c][electrode 0=p,1=n]
    double vb[NR][2];          // Sinv * b (b = surface-row j coupling)
    // per-lane data [field][lane]
#if PLB_BLOCKS_GLOBAL
    double* blk;               // [3][9][LW]: Dinv, Wm, Pm (below)
    double* pad_blk;
#else
    double Dinv[9][LW];        // inverse of the pivoted 3x3 diagonal block D'_x
    double Wm[9][LW];          // W_x = L_x * Dinv_{x-1}     (forward sweep:  y_x = r_x - W_x y_{x-1})
    double Pm[9][LW];          // P_x = Dinv_x * U_x         (backward sweep: u_x = Dinv_x y_x - P_x u_{x+1})
#endif
    double z[3][LW];           // T^{-1} e_I  (border column)
    double q[4][LW];           // j elimination: q_ce, q_pe, q_ps, inv_den
    double jcs[LW];            // a_cs (j row coefficient of the surface concentration)
    double sj[3][LW];          // d(row)/dj for rows ce, pe, ps
#if PLB_SEI
    // aging = :SEI: j, j_s and film are eliminated together node-locally.  Mi = inverse of the 3x3 local
    // matrix (rows j, j_s, film), cpl = couplings of those rows to (c_e, Phi_e, Phi_s) and I:
    // j_ce, j_pe, j_ps, js_pe, js_ps, js_I ; sohc = d(rhs_SOH)/dj_s ; cjv = cj of this factorisation
#if PLB_SEI_LOCAL_GLOBAL
    double sohc[LW];
#else
    double Mi[9][LW], cpl[6][LW], sohc[LW];
#endif
    double cjv, pad2;
#endif
    double schur_inv;          // 1/(g_I - g_ps0*z_ps[0] - g_psN*z_ps[N-1] - g_eta*(z_ps - z_pe)[first anode node])
    double g_ps0, g_psN;
    double g_eta;
    double ionly;              // != 0: the control row has only its I entry (current control): z holds the RAW border
                               // column and dI is folded into the right-hand side before the sweeps (no border solve)
#if PLB_DC
    // concentration-rate control: border row = dc_as * (particle solve, surface) + dc_aj * dj + dc_ace * dc_e, all of lane dc_tgt
    double dc_on, dc_as, dc_aj, dc_ace;
    int dc_tgt, dc_pad;
#endif
};

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
        const double yv[3] = {y0, y1, y2};
        double av[3];
        shfl_from_n<3>(yv, ch.pred, av);
        const double a0 = av[0], a1 = av[1], a2 = av[2];
        // three dependent fused multiply-adds per row: this recurrence is the serial spine of every solve
        y0 = fma(-Wm[2], a2, fma(-Wm[1], a1, fma(-Wm[0], a0, r[0])));
        y1 = fma(-Wm[5], a2, fma(-Wm[4], a1, fma(-Wm[3], a0, r[1])));
        y2 = fma(-Wm[8], a2, fma(-Wm[7], a1, fma(-Wm[6], a0, r[2])));
    }
    {   // the meeting node takes both neighbours
        const int mid = Nx / 2;
        const double yv[3] = {y0, y1, y2};
        double av[3], bv[3];
        shfl_from_n<3>(yv, mid - 1, av);
        shfl_from_n<3>(yv, mid + 1, bv);
        const double a0 = av[0], a1 = av[1], a2 = av[2], b0 = bv[0], b1 = bv[1], b2 = bv[2];
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
        const double uv[3] = {u0, u1, u2};
        double av[3];
        shfl_from_n<3>(uv, ch.succ, av);
        const double a0 = av[0], a1 = av[1], a2 = av[2];
        u0 = fma(-p2, a2, fma(-p1, a1, fma(-p0, a0, c0)));
        u1 = fma(-p5, a2, fma(-p4, a1, fma(-p3, a0, c1)));
        u2 = fma(-p8, a2, fma(-p7, a1, fma(-p6, a0, c2)));
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
        // The particle block is kap*M - cj*I with the constant stencil M = EV diag(EL) EVI (real spectrum,
        // laws_generated.cuh), so its inverse is EV diag(1/(kap*EL_i - cj)) EVI: no elimination, no pivot rounds.
        // Lanes 0-9 of the (first) warp form the ten columns of the cathode block, lanes 16-25 those of the anode.
        const double kap_p = shfl_from(J.kap, 0), kap_n = shfl_from(J.kap, m.Nx - 1);
        const double csj_p = shfl_from(J.cs_j, 0), csj_n = shfl_from(J.cs_j, m.Nx - 1);
        if (lane < 32) {
            const int el = lane >> 4;
            const int c = lane & 15, cc = c < NR ? c : 0;
            const double kap = el == 0 ? kap_p : kap_n;
            const double pd_own = 1.0 / (kap * laws::EL[cc] - cj);
            double t[NR];
#if PLB_SPECTRAL
            // b = cs_j * GJ is a full vector: one more lane per electrode (c == NR) forms vb = Sinv b = EV diag(pd) (EVI GJ) cs_j
            static_assert(NR < 16, "the vb lane of the spectral build");
            const bool vbl = c == NR;
#pragma unroll
            for (int i = 0; i < NR; i++) t[i] = __shfl_sync(FULL, pd_own, (el << 4) + i) * (vbl ? laws::EVIG[i] : laws::EVI[i][cc]);
            if (c <= NR) {
                const double csj = el == 0 ? csj_p : csj_n;
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    double acc = 0.0;
#pragma unroll
                    for (int i = 0; i < NR; i++) acc = fma(laws::EV[r][i], t[i], acc);
                    if (vbl) Fa.vb[r][el] = acc * csj;
                    else Fa.Sinv[r * NR + c][el] = acc;
                }
            }
#else
#pragma unroll
            for (int i = 0; i < NR; i++) t[i] = __shfl_sync(FULL, pd_own, (el << 4) + i) * laws::EVI[i][cc];
            if (c < NR) {
                const double csj = el == 0 ? csj_p : csj_n;
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    double acc = 0.0;
#pragma unroll
                    for (int i = 0; i < NR; i++) acc = fma(laws::EV[r][i], t[i], acc);
                    Fa.Sinv[r * NR + c][el] = acc;
                    if (c == NR - 1) Fa.vb[r][el] = acc * csj;   // vb = Sinv * b,  b = cs_j * e_surf
                }
            }
#endif
        }
        grp_sync();
    }
    // ---- 2. node-local elimination of c_s and j -------------------------------------------------
    const int el = ro.sec == 2 ? 1 : 0;
    double q_ce = 0.0, q_pe = 0.0, q_ps = 0.0, inv_den = 0.0;
#if PLB_SEI
    double qI = 0.0;     // (dj + dj_s) also feels dI through the side-reaction rate law
    {
        // local 3x3 system of (dj, dj_s, dfilm) after the particle elimination (dcs_surf = p0 + p1*dj):
        //   [ j_j + j_cs p1   0       j_film  ] [dj   ]   [ g_j - j_cs p0 - (j_ce, j_pe, j_ps).u          ]
        //   [ js_j            js_js   js_film ] [dj_s ] = [ g_js - (js_pe, js_ps).(u_pe, u_ps) - js_I dI  ]
        //   [ 0               film_js -cj     ] [dfilm]   [ g_film                                        ]
        // cathode lanes: rows/columns 2 and 3 are the identity.  alg_only freezes the film.
        const bool sn = ro.sec == 2;
        const double p1 = alg_only ? 0.0 : -Fa.vb[NR - 1][el];
        double M3[9], Mi[9];
        M3[0] = (ro.elec ? J.j_j : -1.0) + (ro.elec ? J.j_cs * p1 : 0.0); M3[1] = 0.0; M3[2] = (sn && !alg_only) ? J.j_film : 0.0;
        M3[3] = sn ? J.js_j : 0.0; M3[4] = sn ? J.js_js : 1.0; M3[5] = (sn && !alg_only) ? J.js_film : 0.0;
        M3[6] = 0.0; M3[7] = (sn && !alg_only) ? J.film_js : 0.0; M3[8] = (sn && !alg_only) ? -cj : 1.0;
        inv3x3(M3, Mi);
#pragma unroll
        for (int k = 0; k < 9; k++) FA_MI(k) = Mi[k];
        const double c_jce = (ro.elec && !alg_only) ? J.j_ce : 0.0, c_jpe = ro.elec ? J.j_pe : 0.0, c_jps = ro.elec ? J.j_ps : 0.0;
        const double c_spe = sn ? J.js_pe : 0.0, c_sps = sn ? J.js_ps : 0.0, c_sI = sn ? J.js_I : 0.0;
        FA_CPL(0) = c_jce; FA_CPL(1) = c_jpe; FA_CPL(2) = c_jps;
        FA_CPL(3) = c_spe; FA_CPL(4) = c_sps; FA_CPL(5) = c_sI;
        Fa.sohc[lane] = (sn && !alg_only) ? J.soh_js : 0.0;
        if (lane == 0) Fa.cjv = alg_only ? 0.0 : cj;
        // the other rows see dj + dj_s = m.(local right-hand side), m = (1,1,0) Mi
        const double m0 = Mi[0] + Mi[3], m1 = Mi[1] + Mi[4];
        if (ro.elec) {
            q_ce = -m0 * c_jce;
            q_pe = -(m0 * c_jpe + m1 * c_spe);
            q_ps = -(m0 * c_jps + m1 * c_sps);
            qI = -m1 * c_sI;
            inv_den = m0;
        }
    }
#else
    if (ro.elec) {
        // dcs_surf = p0 + p1*dj with p1 = -(Sinv b)[surf]
        const double p1 = alg_only ? 0.0 : -Fa.vb[NR - 1][el];
        inv_den = 1.0 / (-1.0 + J.j_cs * p1);
        q_ce = alg_only ? 0.0 : -J.j_ce * inv_den;
        q_pe = -J.j_pe * inv_den;
        q_ps = -J.j_ps * inv_den;
    }
#endif
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
        shfl_from_n<9>(Di, ch.pred, G);
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
        shfl_from_n<9>(Di, mid + 1, G);
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
#if PLB_BLOCKS_GLOBAL
    {
        double* const gb = Fa.blk;
#pragma unroll
        for (int k = 0; k < 9; k++) { gb[k * LW + lane] = Di[k]; gb[(9 + k) * LW + lane] = Wm[k]; gb[(18 + k) * LW + lane] = Pm[k]; }
    }
#else
#pragma unroll
    for (int k = 0; k < 9; k++) { Fa.Dinv[k][lane] = Di[k]; Fa.Wm[k][lane] = Wm[k]; Fa.Pm[k][lane] = Pm[k]; }
#endif
    // ---- 4. border: z = T^{-1} e_I, then the Schur complement ------------------------------------
#if PLB_SEI
    // border column dF/dI: the Phi_s end rows, plus (through dj + dj_s) the I-dependence of the j_s rows
    const double zf[3] = {sj0 * qI, sj1 * qI, J.ps_I + sj2 * qI};
#else
    const double zf[3] = {0.0, 0.0, J.ps_I};   // border column e_I restricted to this node (Phi_s rows only)
#endif
    // Current control (the headline workload, and newtons_method! for it): the control row is g_I dI = g_I_rhs, so dI is
    // known before the block system is solved and the border column can be folded into the right-hand side of every
    // solve -- no border solve here.  That solve was a third of the factorisation, and the factorising warp is the one
    // the other five wait for at the tick barrier in 61 % of the ticks.
#if PLB_DC
    const bool ionly = ctrl.g_ps0 == 0.0 && ctrl.g_psN == 0.0 && ctrl.g_eta == 0.0 && !ctrl.dc_on;
#else
    const bool ionly = ctrl.g_ps0 == 0.0 && ctrl.g_psN == 0.0 && ctrl.g_eta == 0.0;
#endif
    if (ionly) {
        Fa.z[0][lane] = zf[0]; Fa.z[1][lane] = zf[1]; Fa.z[2][lane] = zf[2];
        // (a singular diagonal block no longer reaches schur_inv through z: look at the blocks themselves)
        const double chk = Di[0] + Di[4] + Di[8];
        const double bad = grp_max((chk == chk && !isinf(chk)) ? 0.0 : 1.0);
        if (lane == 0) {
            Fa.schur_inv = bad != 0.0 ? NAN : 1.0 / ctrl.g_I;
            Fa.g_ps0 = 0.0; Fa.g_psN = 0.0; Fa.g_eta = 0.0; Fa.ionly = 1.0;
#if PLB_DC
            Fa.dc_on = 0.0;
#endif
        }
        grp_sync();
        return;
    }
    double u3[3];
    thomas_sweeps(m.Nx, ch, Wm, Pm, Di, zf, u3);
    Fa.z[0][lane] = u3[0]; Fa.z[1][lane] = u3[1]; Fa.z[2][lane] = u3[2];
    const double z0 = shfl_from(u3[2], 0), zN = shfl_from(u3[2], m.Nx - 1);
    const double ze = shfl_from(u3[2] - u3[1], m.Np + m.Ns);
    double gz_dc = 0.0;
#if PLB_DC
    {
        // the row in terms of the reduced unknowns of lane dc_tgt.  DAE form: -cj on the state itself
        //   surface c_s: d c_s,surf = (S^-1 g_cs)_surf - vb_surf dj      c_e: the block unknown itself
        // newtons_method! form: -d rhs[ind] / d j  (the only algebraic unknown those rows see)
        const int elt = ctrl.dc_tgt >= m.Np + m.Ns ? 1 : 0;
        double as = 0.0, aj = 0.0, ace = 0.0;
        if (ctrl.dc_on) {
            if (ctrl.dc_alg) aj = -shfl_from(ctrl.dc_comp ? J.ce_j : J.cs_j, ctrl.dc_tgt);
            else if (ctrl.dc_comp) ace = -cj;
            else { as = -cj; aj = cj * Fa.vb[NR - 1][elt]; }
        }
        const double djz = ro.elec ? q_ce * u3[0] + q_pe * u3[1] + q_ps * u3[2] : 0.0;      // column solution: no local part
        gz_dc = shfl_from(aj * djz + ace * u3[0], ctrl.dc_tgt);
        if (lane == 0) { Fa.dc_on = ctrl.dc_on ? 1.0 : 0.0; Fa.dc_as = as; Fa.dc_aj = aj; Fa.dc_ace = ace; Fa.dc_tgt = ctrl.dc_tgt; }
    }
#endif
    if (lane == 0) {
        Fa.schur_inv = 1.0 / (ctrl.g_I - ctrl.g_ps0 * z0 - ctrl.g_psN * zN - ctrl.g_eta * ze - gz_dc);
        Fa.g_ps0 = ctrl.g_ps0;
        Fa.g_psN = ctrl.g_psN;
        Fa.g_eta = ctrl.g_eta;
        Fa.ionly = 0.0;
    }
    grp_sync();
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
#if PLB_SEI
    const bool sn = ro.sec == 2;
    double Mi[9], cplr[6];
#pragma unroll
    for (int k = 0; k < 9; k++) Mi[k] = FA_MI(k);
#pragma unroll
    for (int k = 0; k < 6; k++) cplr[k] = FA_CPL(k);
    // local right-hand side of (j, j_s, film) after the particle elimination
    double v0 = ro.elec ? g.j - Fa.jcs[lane] * p0 : 0.0;
    double v1 = sn ? g.js : 0.0;
    const double v2 = (sn && !alg_only) ? g.film : 0.0;
    const double q0 = ro.elec ? (Mi[0] + Mi[3]) * v0 + (Mi[1] + Mi[4]) * v1 + (Mi[2] + Mi[5]) * v2 : 0.0;
#else
    const double inv_den = Fa.q[3][lane];
    const double q0 = ro.elec ? (g.j - Fa.jcs[lane] * p0) * inv_den : 0.0;
#endif
    double rf[3];
    rf[0] = alg_only ? 0.0 : g.ce - Fa.sj[0][lane] * q0;
    rf[1] = g.pe - Fa.sj[1][lane] * q0;
    rf[2] = (ro.elec ? g.ps : 0.0) - Fa.sj[2][lane] * q0;
    if (!ro.act) { rf[0] = rf[1] = rf[2] = 0.0; }
    double Di[9], Wm[9], Pm[9];
#if PLB_BLOCKS_GLOBAL
    {
        const double* const gb = Fa.blk;
#pragma unroll
        for (int k = 0; k < 9; k++) { Di[k] = gb[k * LW + lane]; Wm[k] = gb[(9 + k) * LW + lane]; Pm[k] = gb[(18 + k) * LW + lane]; }
    }
#else
#pragma unroll
    for (int k = 0; k < 9; k++) { Di[k] = Fa.Dinv[k][lane]; Wm[k] = Fa.Wm[k][lane]; Pm[k] = Fa.Pm[k][lane]; }
#endif
    double u3[3];
    const LaneChain ch = make_chain(m.Nx, lane);
    double dI;
    if (Fa.ionly != 0.0) {       // (uniform) current control: dI first, its column folded into the right-hand side
        dI = gI * Fa.schur_inv;
        rf[0] -= Fa.z[0][lane] * dI; rf[1] -= Fa.z[1][lane] * dI; rf[2] -= Fa.z[2][lane] * dI;
        thomas_sweeps(m.Nx, ch, Wm, Pm, Di, rf, u3);
    } else {
        thomas_sweeps(m.Nx, ch, Wm, Pm, Di, rf, u3);
        // border
        const double x0 = shfl_from(u3[2], 0), xN = shfl_from(u3[2], m.Nx - 1);
        double gx = Fa.g_ps0 * x0 + Fa.g_psN * xN;
        if (Fa.g_eta != 0.0) gx += Fa.g_eta * shfl_from(u3[2] - u3[1], m.Np + m.Ns);      // eta_p control (uniform branch)
#if PLB_DC
        if (Fa.dc_on != 0.0) {                                                              // (uniform branch)
            const double dj0 = ro.elec ? q0 + Fa.q[0][lane] * u3[0] + Fa.q[1][lane] * u3[1] + Fa.q[2][lane] * u3[2] : 0.0;
            gx += shfl_from(Fa.dc_as * p0 + Fa.dc_aj * dj0 + Fa.dc_ace * u3[0], Fa.dc_tgt);
        }
#endif
        dI = (gI - gx) * Fa.schur_inv;
        u3[0] -= Fa.z[0][lane] * dI; u3[1] -= Fa.z[1][lane] * dI; u3[2] -= Fa.z[2][lane] * dI;
    }
    // back-substitute j (and j_s, film) and the particle
#if PLB_SEI
    v0 -= cplr[0] * u3[0] + cplr[1] * u3[1] + cplr[2] * u3[2];
    v1 -= cplr[3] * u3[1] + cplr[4] * u3[2] + cplr[5] * dI;
    const double dj = ro.elec ? Mi[0] * v0 + Mi[1] * v1 + Mi[2] * v2 : 0.0;
    const double djs = sn ? Mi[3] * v0 + Mi[4] * v1 + Mi[5] * v2 : 0.0;
    const double dfilm = (sn && !alg_only) ? Mi[6] * v0 + Mi[7] * v1 + Mi[8] * v2 : 0.0;
    // SOH: its column only holds -cj on the diagonal; its row is dense over j_s (residuals.jl:278-297)
    const double sj_sum = warp_sum(Fa.sohc[lane] * djs);
    g.soh = alg_only ? 0.0 : (g.soh - sj_sum) / (-Fa.cjv);
    g.js = djs;
    g.film = dfilm;
#else
    const double dj = ro.elec ? q0 + Fa.q[0][lane] * u3[0] + Fa.q[1][lane] * u3[1] + Fa.q[2][lane] * u3[2] : 0.0;
#endif
    g.ce = alg_only ? 0.0 : u3[0];
    g.pe = u3[1];
    g.ps = ro.elec ? u3[2] : 0.0;
    g.j = dj;
#pragma unroll
    for (int r = 0; r < NR; r++) g.cs[r] = (ro.elec && !alg_only) ? s[r] - Fa.vb[r][el] * dj : 0.0;
    return dI;
}

#else   // PLB_TH ===================================================================================
// ------------------------------------------------------------------------------------------------
// structured Newton-matrix factorisation / solve, thermal variant (N = 351).  The algebra below is
// the one prototyped (and checked against dense LAPACK solves of the oracle's Jacobian) in
// tests/proto_thermal_solver.py:
//   1. particle block kap(T_x)*MC - cj I differs from node to node (Arrhenius D_s), but MC = EV diag(EL) EVI
//      is a fixed matrix with a real spectrum, so the block is diagonal in a constant basis: a particle
//      solve is two constant 10x10 mat-vecs and ten reciprocals kept per node (no per-node LU);
//   2. c_s and j are eliminated node-locally (T-row included: it sees j and the surface concentration);
//   3. the current-collector temperature chains (scalar tridiagonal, one node per lane) are folded into
//      the T-diagonal of the two end nodes;
//   4. 4x4-block tridiagonal system in (c_e, Phi_e, Phi_s, T): twisted block-Thomas along the lanes.
//      The T rows of four nodes (the ends of the cell and the inner ends of the electrodes) reach two
//      nodes away through the one-sided differences of thermal_derivatives; with the twisted ordering
//      those couplings point either two nodes back (absorbed into the incoming block and the right-hand
//      side) or, at the chain heads, two nodes ahead (absorbed into the outgoing block of the next node
//      and into the last back-substitution step) -- exact, no fill outside the block tridiagonal;
//   5. the applied current is a border, as in the isothermal variant.
// ------------------------------------------------------------------------------------------------
// The factored 4x4 blocks (Dinv, W, P: 48 doubles per lane, 12 KB per system) live in the system's slot of the GLOBAL workspace
// (L2-resident), not in shared memory: they are written once per factorisation and read once at the head of a solve (48
// independent coalesced loads per lane), and those 12 KB were what held the thermal families at 5 / 3 systems per SM.
#ifndef PLB_TH_BLOCKS_GLOBAL
#define PLB_TH_BLOCKS_GLOBAL 1
#endif
// ... and so do the particle reciprocals pd and the T column of the particle rows in the eigen-basis wT (2 N_r doubles per lane,
// read twice per solve): 5 KB per system more
#ifndef PLB_TH_PD_GLOBAL
#define PLB_TH_PD_GLOBAL PLB_TH_BLOCKS_GLOBAL
#endif
constexpr int FA_GLOBAL = PLB_TH_BLOCKS_GLOBAL ? (48 + (PLB_TH_PD_GLOBAL ? 2 * NR : 0)) * LW : 0;      // doubles per system slot of the global workspace
#if PLB_TH_PD_GLOBAL
#define FA_PD(i) Fa.blk[(48 + (i)) * LW + lane]
#define FA_WT(i) Fa.blk[(48 + NR + (i)) * LW + lane]
#else
#define FA_PD(i) Fa.pd[i][lane]
#define FA_WT(i) Fa.wT[i][lane]
#endif
struct WarpFactor {
#if PLB_TH_BLOCKS_GLOBAL
    double* blk;               // [3][16][LW]: Dinv, Wm, Pm
    double* pad_blk;
#else
    double Dinv[16][LW], Wm[16][LW], Pm[16][LW];
#endif
    double Fr[4][LW];          // T-row multiplier for the node two behind in the chain
    double Eo[3][LW];          // chain heads: T-row coupling to (c_e, Phi_e, Phi_s) two nodes ahead
    double z[4][LW], zx[LW];   // border column solution
    double q[5][LW];           // j elimination: q_ce, q_pe, q_ps, q_T, inv_den
    double jcs[LW];
    double sj[4][LW];          // effective d(row)/dj for rows ce, pe, ps, T
    double tcs[LW];            // T-row coefficient of the surface concentration
#if !PLB_TH_PD_GLOBAL
    double pd[NR][LW];         // 1 / (kap_x * EL_i - cj)
    double wT[NR][LW];         // EVI * (d res_cs / dT)
#endif
    double csj[LW];
    double chm[LW], chip[LW], chup[LW], hm[LW];   // collector chains: multiplier, 1/pivot, successor coupling; end-node multiplier
    double gT[LW], gX[LW];     // dT control: border-row entries on this lane's T / collector T
#if PLB_SEI
    // aging = :SEI: (j, j_s, film) are eliminated together node-locally (cathode / separator lanes: identity rows).
    //   M l + Cl u + cI dI = v   ->   l = Mi v + Ql u + QI dI,  Ql = -Mi Cl (3x4), QI = -Mi cI
    //   reduced rows (c_e, Phi_e, Phi_s, T):  D += Sj Ql,  r -= SM v (SM = Sj Mi, 4x3),  border column += Sj QI
    double Mi[9][LW], Ql[12][LW], QI[3][LW], SM[12][LW];
    double sohc[LW];           // d(rhs_SOH)/dj_s
    double pdjs[LW];           // dT control inside newtons_method!: border-row coefficient of dj_s
    double cjv, pad4;
#endif
    double schur_inv, g_ps0, g_psN;
    double mode;               // border row: 0 (Phi_s ends + I), 1 dT in the DAE, 2 dT inside newtons_method!
    double g_eta, pad3;        // eta_p control
};

// border row times a block solution.  mode 1: the row has entries on every temperature; mode 2 (algebraic
// initialisation of the dT mode): the row is -sum_x w_x d(rhs_T[x])/d(j, Phi_e, Phi_s), whose coefficients
// are parked in Fa.pd[0..9] / Fa.wT[0] (the particle data is not used in that mode):
//   pd[0]: j ; pd[1..5]: Phi_e at x-2..x+2 ; pd[6..9], wT[0]: Phi_s at x-2..x+2
__device__ __forceinline__ double border_dot(const ModelDesc& m, const WarpFactor& Fa, int mode, const double* u4,
                                             double ux, double dj, double djs, int lane) {
    const double x0 = shfl_from(u4[2], 0), xN = shfl_from(u4[2], m.Nx - 1);
    double g = Fa.g_ps0 * x0 + Fa.g_psN * xN;
    if (Fa.g_eta != 0.0) g += Fa.g_eta * shfl_from(u4[2] - u4[1], m.Np + m.Ns);
    if (mode == 1) g += warp_sum(Fa.gT[lane] * u4[3] + Fa.gX[lane] * ux);
    if (mode == 2) {
        double a = FA_PD(0) * dj;
#if PLB_SEI
        a = fma(Fa.pdjs[lane], djs, a);
#else
        (void)djs;
#endif
        a = fma(FA_PD(1), shfl_up2(u4[1]), a);
        a = fma(FA_PD(2), shfl_up(u4[1]), a);
        a = fma(FA_PD(3), u4[1], a);
        a = fma(FA_PD(4), shfl_dn(u4[1]), a);
        a = fma(FA_PD(5), shfl_dn2(u4[1]), a);
        a = fma(FA_PD(6), shfl_up2(u4[2]), a);
        a = fma(FA_PD(7), shfl_up(u4[2]), a);
        a = fma(FA_PD(8), u4[2], a);
        a = fma(FA_PD(9), shfl_dn(u4[2]), a);
        a = fma(FA_WT(0), shfl_dn2(u4[2]), a);
        g += warp_sum(a);
    }
    return g;
}

// 4x4 inverse by the adjugate (2x2 sub-determinants), row-major
__device__ __forceinline__ void inv4x4(const double* a, double* b) {
    const double s0 = a[0] * a[5] - a[4] * a[1], s1 = a[0] * a[6] - a[4] * a[2], s2 = a[0] * a[7] - a[4] * a[3];
    const double s3 = a[1] * a[6] - a[5] * a[2], s4 = a[1] * a[7] - a[5] * a[3], s5 = a[2] * a[7] - a[6] * a[3];
    const double c5 = a[10] * a[15] - a[14] * a[11], c4 = a[9] * a[15] - a[13] * a[11], c3 = a[9] * a[14] - a[13] * a[10];
    const double c2 = a[8] * a[15] - a[12] * a[11], c1 = a[8] * a[14] - a[12] * a[10], c0 = a[8] * a[13] - a[12] * a[9];
    const double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    const double i = 1.0 / det;
    b[0] = (a[5] * c5 - a[6] * c4 + a[7] * c3) * i;
    b[1] = (-a[1] * c5 + a[2] * c4 - a[3] * c3) * i;
    b[2] = (a[13] * s5 - a[14] * s4 + a[15] * s3) * i;
    b[3] = (-a[9] * s5 + a[10] * s4 - a[11] * s3) * i;
    b[4] = (-a[4] * c5 + a[6] * c2 - a[7] * c1) * i;
    b[5] = (a[0] * c5 - a[2] * c2 + a[3] * c1) * i;
    b[6] = (-a[12] * s5 + a[14] * s2 - a[15] * s1) * i;
    b[7] = (a[8] * s5 - a[10] * s2 + a[11] * s1) * i;
    b[8] = (a[4] * c4 - a[5] * c2 + a[7] * c0) * i;
    b[9] = (-a[0] * c4 + a[1] * c2 - a[3] * c0) * i;
    b[10] = (a[12] * s4 - a[13] * s2 + a[15] * s0) * i;
    b[11] = (-a[8] * s4 + a[9] * s2 - a[11] * s0) * i;
    b[12] = (-a[4] * c3 + a[5] * c1 - a[6] * c0) * i;
    b[13] = (a[0] * c3 - a[1] * c1 + a[2] * c0) * i;
    b[14] = (-a[12] * s3 + a[13] * s1 - a[14] * s0) * i;
    b[15] = (a[8] * s3 - a[9] * s1 + a[10] * s0) * i;
}

// chain bookkeeping of one lane for the twisted elimination (meeting node m.mid, inside the separator)
struct LaneChain {
    int pred, succ, pred2, succ2;   // one / two nodes back and ahead along this lane's chain
    int pos;                        // position in the chain (heads 0), -1 for the meeting node / inactive lanes
    bool is_mid, rightc;
    int n_in, n_out;                // iterations of the inward / outward sweeps
    int posA, posB;                 // chain positions of the two inner electrode ends (T rows reaching two back)
    // current-collector chains
    int cpred, csucc, ctail;        // chain neighbours; ctail: lane owning the chain node adjacent to this END node
    bool ch, is_tail;
    int nch;
};
__device__ __forceinline__ LaneChain make_chain(const ModelDesc& m, const LaneRole& ro, int lane) {
    LaneChain c;
    const int Nx = m.Nx, mid = m.mid;
    c.is_mid = lane == mid;
    const bool left = lane < mid, right = lane > mid && lane < Nx;
    c.rightc = right;
    c.pred = left ? (lane > 0 ? lane - 1 : lane) : (right ? (lane < Nx - 1 ? lane + 1 : lane) : (lane == mid ? lane - 1 : lane));
    c.succ = left ? lane + 1 : (right ? lane - 1 : lane);
    c.pred2 = left ? (lane > 1 ? lane - 2 : 0) : (right ? (lane < Nx - 2 ? lane + 2 : Nx - 1) : lane);
    c.succ2 = left ? (lane + 2 < Nx ? lane + 2 : Nx - 1) : (right ? (lane >= 2 ? lane - 2 : 0) : lane);
    c.pos = left ? lane : (right ? Nx - 1 - lane : -1);
    const int a = mid - 1, b = Nx - mid - 2;
    c.n_in = a > b ? a : b;
    c.n_out = mid > Nx - 1 - mid ? mid : Nx - 1 - mid;
    c.posA = m.Np - 1;
    c.posB = m.Nn - 1;
    c.ch = ro.cha || ro.chz;
    c.cpred = ro.cha ? (lane > 0 ? lane - 1 : 0) : (lane < LW - 1 ? lane + 1 : LW - 1);
    c.csucc = ro.cha ? (lane < LW - 1 ? lane + 1 : LW - 1) : (lane > 0 ? lane - 1 : 0);
    c.is_tail = (ro.cha && lane == m.Na - 1) || (ro.chz && lane == Nx - m.Nz);
    c.ctail = lane == 0 ? m.Na - 1 : Nx - m.Nz;
    c.nch = m.Na > m.Nz ? m.Na : m.Nz;
    return c;
}

// dense 4x4 (row-major) times 4-vector
#define PLB_MV4(M, v0, v1, v2, v3, r)                                                  \
    ((M)[(r) * 4 + 0] * (v0) + (M)[(r) * 4 + 1] * (v1) + (M)[(r) * 4 + 2] * (v2) + (M)[(r) * 4 + 3] * (v3))

// Solve the matrix WITHOUT the border: block right-hand side rb[4] (after the node-local elimination),
// chain right-hand side rx; returns u[4], ux.  Di/Wm/Pm: this lane's factor blocks in registers.
__device__ __forceinline__ void core_solve(const ModelDesc& m, const LaneChain& ch, const WarpFactor& Fa,
                                           const double* Di, const double* Wm, const double* Pm,
                                           const double* rb, double rx, double* u, double& ux, int lane) {
    // 1. collector chains, forward
    const double chm = Fa.chm[lane];
    double yx = rx;
#pragma unroll 1
    for (int it = 0; it < ch.nch - 1; it++) yx = rx - chm * shfl_from(yx, ch.cpred);
    // 2. fold into the two end nodes
    double r0 = rb[0], r1 = rb[1], r2 = rb[2], r3 = rb[3] - Fa.hm[lane] * shfl_from(yx, ch.ctail);
    // 3. twisted block-Thomas, inward
    const double f0 = Fa.Fr[0][lane], f1 = Fa.Fr[1][lane], f2 = Fa.Fr[2][lane], f3 = Fa.Fr[3][lane];
    double y0 = r0, y1 = r1, y2 = r2, y3 = r3;
#pragma unroll 1
    for (int it = 0; it < ch.n_in; it++) {
        if (it == ch.posA - 1 || it == ch.posB - 1) {
            // the T row of an inner electrode end also sees the node two back, final by now
            const double b0 = shfl_from(y0, ch.pred2), b1 = shfl_from(y1, ch.pred2), b2 = shfl_from(y2, ch.pred2), b3 = shfl_from(y3, ch.pred2);
            if (ch.pos == it + 1) r3 -= f0 * b0 + f1 * b1 + f2 * b2 + f3 * b3;
        }
        const double a0 = shfl_from(y0, ch.pred), a1 = shfl_from(y1, ch.pred), a2 = shfl_from(y2, ch.pred), a3 = shfl_from(y3, ch.pred);
        y0 = r0 - PLB_MV4(Wm, a0, a1, a2, a3, 0);
        y1 = r1 - PLB_MV4(Wm, a0, a1, a2, a3, 1);
        y2 = r2 - PLB_MV4(Wm, a0, a1, a2, a3, 2);
        y3 = r3 - PLB_MV4(Wm, a0, a1, a2, a3, 3);
    }
    {   // the meeting node takes both neighbours
        const int mid = m.mid;
        const double a0 = shfl_from(y0, mid - 1), a1 = shfl_from(y1, mid - 1), a2 = shfl_from(y2, mid - 1), a3 = shfl_from(y3, mid - 1);
        const double b0 = shfl_from(y0, mid + 1), b1 = shfl_from(y1, mid + 1), b2 = shfl_from(y2, mid + 1), b3 = shfl_from(y3, mid + 1);
        if (ch.is_mid) {
            y0 = r0 - PLB_MV4(Wm, a0, a1, a2, a3, 0) - PLB_MV4(Pm, b0, b1, b2, b3, 0);
            y1 = r1 - PLB_MV4(Wm, a0, a1, a2, a3, 1) - PLB_MV4(Pm, b0, b1, b2, b3, 1);
            y2 = r2 - PLB_MV4(Wm, a0, a1, a2, a3, 2) - PLB_MV4(Pm, b0, b1, b2, b3, 2);
            y3 = r3 - PLB_MV4(Wm, a0, a1, a2, a3, 3) - PLB_MV4(Pm, b0, b1, b2, b3, 3);
        }
    }
    const double c0 = PLB_MV4(Di, y0, y1, y2, y3, 0), c1 = PLB_MV4(Di, y0, y1, y2, y3, 1);
    const double c2 = PLB_MV4(Di, y0, y1, y2, y3, 2), c3 = PLB_MV4(Di, y0, y1, y2, y3, 3);
    double u0 = c0, u1 = c1, u2 = c2, u3 = c3;
    const double pz = ch.is_mid ? 0.0 : 1.0;   // the meeting node has no outward update
#pragma unroll 1
    for (int it = 0; it < ch.n_out; it++) {
        const double a0 = shfl_from(u0, ch.succ), a1 = shfl_from(u1, ch.succ), a2 = shfl_from(u2, ch.succ), a3 = shfl_from(u3, ch.succ);
        u0 = c0 - pz * PLB_MV4(Pm, a0, a1, a2, a3, 0);
        u1 = c1 - pz * PLB_MV4(Pm, a0, a1, a2, a3, 1);
        u2 = c2 - pz * PLB_MV4(Pm, a0, a1, a2, a3, 2);
        u3 = c3 - pz * PLB_MV4(Pm, a0, a1, a2, a3, 3);
    }
    {   // chain heads: their T row also reaches two nodes ahead
        const double a0 = shfl_from(u0, ch.succ2), a1 = shfl_from(u1, ch.succ2), a2 = shfl_from(u2, ch.succ2);
        const double t = Fa.Eo[0][lane] * a0 + Fa.Eo[1][lane] * a1 + Fa.Eo[2][lane] * a2;
        u0 -= Di[3] * t; u1 -= Di[7] * t; u2 -= Di[11] * t; u3 -= Di[15] * t;
    }
    u[0] = u0; u[1] = u1; u[2] = u2; u[3] = u3;
    // 4. collector chains, backward
    const double uT = shfl_from(u3, lane < m.Na ? 0 : m.Nx - 1);
    const double chip = Fa.chip[lane], chup = Fa.chup[lane];
    double v = yx * chip;
#pragma unroll 1
    for (int it = 0; it < ch.nch; it++) {
        const double us = shfl_from(v, ch.csucc);
        v = (yx - chup * (ch.is_tail ? uT : us)) * chip;
    }
    ux = ch.ch ? v : 0.0;
}

// alg_only: Newton on the algebraic block (newtons_method!, model_evaluation.jl:430-480): c_e, c_s and
// all temperatures are frozen, the differential rows are replaced by identity.
__device__ __forceinline__ void warp_factor_impl(const ModelDesc& m, const LaneRole& ro, const LaneJac& J,
                                                 const CtrlRow& ctrl, double cj, bool alg_only,
                                                 WarpFactor& Fa, int lane) {
    const LaneChain ch = make_chain(m, ro, lane);
    const bool dyn = !alg_only;
    // ---- 1. particles in the eigen-basis of MC ---------------------------------------------------
    double beta = 0.0, tau = 0.0;
    if (dyn) {
        double pd[NR];
#pragma unroll
        for (int i = 0; i < NR; i++) {
            pd[i] = ro.elec ? 1.0 / (J.kap * laws::EL[i] - cj) : 0.0;
            FA_PD(i) = pd[i];
        }
#pragma unroll
        for (int i = 0; i < NR; i++) {
            double w = 0.0;
#pragma unroll
            for (int c = 0; c < NR; c++) w = fma(laws::EVI[i][c], J.csT[c], w);
            FA_WT(i) = w;
            beta = fma(laws::EV[NR - 1][i] * pd[i], PLB_EVIB(i), beta);
            tau = fma(laws::EV[NR - 1][i] * pd[i], w, tau);
        }
        beta *= J.cs_j;
    }
    Fa.csj[lane] = J.cs_j;
    // ---- 2. node-local elimination of c_s and j (and j_s, film) ----------------------------------
#if PLB_SEI
    const bool sn = ro.sec == 2;
    double Ql[12], QIv[3], Sj[12];
    {
        double M3[9], Mi[9];
        M3[0] = ro.elec ? J.j_j - J.j_cs * beta : -1.0; M3[1] = 0.0; M3[2] = (sn && dyn) ? J.j_film : 0.0;
        M3[3] = sn ? J.js_j : 0.0; M3[4] = sn ? J.js_js : 1.0; M3[5] = (sn && dyn) ? J.js_film : 0.0;
        M3[6] = 0.0; M3[7] = (sn && dyn) ? J.film_js : 0.0; M3[8] = (sn && dyn) ? -cj : 1.0;
        inv3x3(M3, Mi);
        const double Cl[12] = {(ro.elec && dyn) ? J.j_ce : 0.0, ro.elec ? J.j_pe : 0.0, ro.elec ? J.j_ps : 0.0,
                               (ro.elec && dyn) ? J.j_T - J.j_cs * tau : 0.0,
                               0.0, sn ? J.js_pe : 0.0, sn ? J.js_ps : 0.0, (sn && dyn) ? J.js_T : 0.0,
                               0.0, 0.0, 0.0, 0.0};
        const double cI[3] = {0.0, sn ? J.js_I : 0.0, 0.0};
        const double Sjv[12] = {(ro.elec && dyn) ? J.ce_j : 0.0, (sn && dyn) ? J.ce_j : 0.0, 0.0,
                                ro.elec ? J.pe_j : 0.0, sn ? J.pe_j : 0.0, 0.0,
                                ro.elec ? J.ps_j : 0.0, sn ? J.ps_j : 0.0, 0.0,
                                (ro.elec && dyn) ? J.T_j - J.T_cs * beta : 0.0, (sn && dyn) ? J.T_js : 0.0, (sn && dyn) ? J.T_film : 0.0};
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
            for (int c = 0; c < 4; c++) Ql[r * 4 + c] = -(Mi[r * 3 + 0] * Cl[c] + Mi[r * 3 + 1] * Cl[4 + c] + Mi[r * 3 + 2] * Cl[8 + c]);
            QIv[r] = -(Mi[r * 3 + 0] * cI[0] + Mi[r * 3 + 1] * cI[1] + Mi[r * 3 + 2] * cI[2]);
        }
#pragma unroll
        for (int k = 0; k < 12; k++) { Sj[k] = Sjv[k]; Fa.Ql[k][lane] = Ql[k]; }
#pragma unroll
        for (int k = 0; k < 9; k++) Fa.Mi[k][lane] = Mi[k];
#pragma unroll
        for (int k = 0; k < 3; k++) Fa.QI[k][lane] = QIv[k];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 3; c++)
                Fa.SM[r * 3 + c][lane] = Sj[r * 3 + 0] * Mi[c] + Sj[r * 3 + 1] * Mi[3 + c] + Sj[r * 3 + 2] * Mi[6 + c];
        Fa.sohc[lane] = (sn && dyn) ? J.soh_js : 0.0;
        if (lane == 0) Fa.cjv = dyn ? cj : 0.0;
    }
#else
    double q[4] = {0.0, 0.0, 0.0, 0.0}, inv_den = 0.0;
    if (ro.elec) {
        inv_den = 1.0 / (-1.0 - J.j_cs * beta);
        q[0] = dyn ? -J.j_ce * inv_den : 0.0;
        q[1] = -J.j_pe * inv_den;
        q[2] = -J.j_ps * inv_den;
        q[3] = dyn ? -(J.j_T - J.j_cs * tau) * inv_den : 0.0;
    }
    const double sj[4] = {dyn ? J.ce_j : 0.0, J.pe_j, J.ps_j, dyn ? J.T_j - J.T_cs * beta : 0.0};
#pragma unroll
    for (int k = 0; k < 4; k++) { Fa.q[k][lane] = q[k]; Fa.sj[k][lane] = sj[k]; }
    Fa.q[4][lane] = inv_den;
#endif
    Fa.jcs[lane] = J.j_cs;
    Fa.tcs[lane] = dyn ? J.T_cs : 0.0;
    double Dm[16];
    Dm[0] = dyn ? J.ceD - cj : 1.0; Dm[1] = 0.0; Dm[2] = 0.0; Dm[3] = 0.0;
    Dm[4] = dyn ? J.pcD : 0.0; Dm[5] = J.peD; Dm[6] = 0.0; Dm[7] = dyn ? J.peTD : 0.0;
    Dm[8] = 0.0; Dm[9] = 0.0; Dm[10] = J.psD; Dm[11] = 0.0;
    Dm[12] = dyn ? J.T_ce[2] : 0.0; Dm[13] = dyn ? J.T_pe[2] : 0.0; Dm[14] = dyn ? J.T_ps[2] : 0.0;
    Dm[15] = dyn ? J.T_TD - cj - J.T_cs * tau : 1.0;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
#if PLB_SEI
            Dm[r * 4 + c] += Sj[r * 3 + 0] * Ql[c] + Sj[r * 3 + 1] * Ql[4 + c] + Sj[r * 3 + 2] * Ql[8 + c];
#else
            Dm[r * 4 + c] = fma(sj[r], q[c], Dm[r * 4 + c]);
#endif
        }
    if (!ro.act) {
#pragma unroll
        for (int k = 0; k < 16; k++) Dm[k] = (k % 5 == 0) ? 1.0 : 0.0;
    }
    // off-diagonal blocks, 9 structural entries each: (0,0) (1,0) (1,1) (1,3) (2,2) (3,0) (3,1) (3,2) (3,3)
    double L9[9] = {dyn ? J.ceL : 0.0, dyn ? J.pcL : 0.0, J.peL, dyn ? J.peTL : 0.0, J.psL,
                    dyn ? J.T_ce[1] : 0.0, dyn ? J.T_pe[1] : 0.0, dyn ? J.T_ps[1] : 0.0, dyn ? J.T_TL : 0.0};
    double U9[9] = {dyn ? J.ceU : 0.0, dyn ? J.pcU : 0.0, J.peU, dyn ? J.peTU : 0.0, J.psU,
                    dyn ? J.T_ce[3] : 0.0, dyn ? J.T_pe[3] : 0.0, dyn ? J.T_ps[3] : 0.0, dyn ? J.T_TU : 0.0};
    if (!ro.act || ro.x == 0) {
#pragma unroll
        for (int k = 0; k < 9; k++) L9[k] = 0.0;
    }
    if (!ro.act || ro.x >= m.Nx - 1) {
#pragma unroll
        for (int k = 0; k < 9; k++) U9[k] = 0.0;
    }
    const bool useE = dyn && ro.act;
    const double E2m[3] = {useE ? J.T_ce[0] : 0.0, useE ? J.T_pe[0] : 0.0, useE ? J.T_ps[0] : 0.0};
    const double E2p[3] = {useE ? J.T_ce[4] : 0.0, useE ? J.T_pe[4] : 0.0, useE ? J.T_ps[4] : 0.0};
    // ---- 3. collector chains ----------------------------------------------------------------------
    {
        const bool chd = ch.ch && dyn;
        const double cin = chd ? (ro.cha ? J.Tx_L : J.Tx_U) : 0.0;
        const double cout = chd ? (ro.cha ? J.Tx_U : J.Tx_L) : 0.0;
        const double di = chd ? J.Tx_D - cj : 1.0;
        const double cop = shfl_from(cout, ch.cpred);
        double pv = di, mm = 0.0;
#pragma unroll 1
        for (int it = 0; it < ch.nch - 1; it++) {
            const double pp = shfl_from(pv, ch.cpred);
            mm = cin / pp;
            pv = di - mm * cop;
        }
        Fa.chm[lane] = mm; Fa.chip[lane] = 1.0 / pv; Fa.chup[lane] = cout;
        const double pvt = shfl_from(pv, ch.ctail), cot = shfl_from(cout, ch.ctail);
        const double hc = !dyn ? 0.0 : (ro.x == 0 ? J.T_TL : (ro.x == m.Nx - 1 ? J.T_TU : 0.0));
        const double hm = hc / pvt;
        Dm[15] -= hm * cot;
        Fa.hm[lane] = hm;
    }
    // ---- 4. twisted block-Thomas factorisation ------------------------------------------------------
    double Cin[9], Cout[9], Cp[9], Ein[3], Eout[3];
#pragma unroll
    for (int k = 0; k < 9; k++) { Cin[k] = ch.rightc ? U9[k] : L9[k]; Cout[k] = ch.rightc ? L9[k] : U9[k]; }
#pragma unroll
    for (int k = 0; k < 3; k++) { Ein[k] = ch.rightc ? E2p[k] : E2m[k]; Eout[k] = ch.rightc ? E2m[k] : E2p[k]; }
    const bool has_pred = ch.pred != lane;
    if (!has_pred) {
#pragma unroll
        for (int k = 0; k < 9; k++) Cin[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) Cp[k] = shfl_from(Cout[k], ch.pred);
    const bool hasF = ch.pos >= 2 && (Ein[0] != 0.0 || Ein[1] != 0.0 || Ein[2] != 0.0);
    double Di[16], Wm[16], Fr[4] = {0.0, 0.0, 0.0, 0.0};
    double Xo[12], Xp[12];   // dense 4x3 extension of the outgoing block of a position-1 node, and the predecessor's
#pragma unroll
    for (int k = 0; k < 12; k++) { Xo[k] = 0.0; Xp[k] = 0.0; }
    inv4x4(Dm, Di);
#pragma unroll
    for (int k = 0; k < 16; k++) Wm[k] = 0.0;
#pragma unroll 1
    for (int it = 0; it < ch.n_in + 1; it++) {
        double G[16];
#pragma unroll
        for (int k = 0; k < 16; k++) G[k] = shfl_from(Di[k], ch.pred);
        if (it == ch.posA - 1 || it == ch.posB - 1) {
            // absorb the coupling to the node two back: F = Ein * Dinv_{p-2}; Cin(T row) -= F * Cout_{p-2}
            double Fn[4] = {0.0, 0.0, 0.0, 0.0}, Cq[9];
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int c = 0; c < 4; c++) Fn[c] = fma(Ein[k], shfl_from(Di[k * 4 + c], ch.pred2), Fn[c]);
#pragma unroll
            for (int k = 0; k < 9; k++) Cq[k] = shfl_from(Cout[k], ch.pred2);
            if (ch.pos == it + 1 && hasF) {
#pragma unroll
                for (int c = 0; c < 4; c++) Fr[c] = Fn[c];
                Cin[5] -= Fn[0] * Cq[0] + Fn[1] * Cq[1] + Fn[3] * Cq[5];
                Cin[6] -= Fn[1] * Cq[2] + Fn[3] * Cq[6];
                Cin[7] -= Fn[2] * Cq[4] + Fn[3] * Cq[7];
                Cin[8] -= Fn[1] * Cq[3] + Fn[3] * Cq[8];
            }
        }
        // W = Cin * G (Cin sparse)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            Wm[c] = Cin[0] * G[c];
            Wm[4 + c] = Cin[1] * G[c] + Cin[2] * G[4 + c] + Cin[3] * G[12 + c];
            Wm[8 + c] = Cin[4] * G[8 + c];
            Wm[12 + c] = Cin[5] * G[c] + Cin[6] * G[4 + c] + Cin[7] * G[8 + c] + Cin[8] * G[12 + c];
        }
        if (it == 1 && ch.pos == 2) {
            // the predecessor's outgoing block was extended by a chain head: fold W * Xp into D for good
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 3; c++)
                    Dm[r * 4 + c] -= Wm[r * 4 + 0] * Xp[c] + Wm[r * 4 + 1] * Xp[3 + c] + Wm[r * 4 + 2] * Xp[6 + c] + Wm[r * 4 + 3] * Xp[9 + c];
        }
        double Dp[16];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            Dp[r * 4 + 0] = Dm[r * 4 + 0] - (Wm[r * 4 + 0] * Cp[0] + Wm[r * 4 + 1] * Cp[1] + Wm[r * 4 + 3] * Cp[5]);
            Dp[r * 4 + 1] = Dm[r * 4 + 1] - (Wm[r * 4 + 1] * Cp[2] + Wm[r * 4 + 3] * Cp[6]);
            Dp[r * 4 + 2] = Dm[r * 4 + 2] - (Wm[r * 4 + 2] * Cp[4] + Wm[r * 4 + 3] * Cp[7]);
            Dp[r * 4 + 3] = Dm[r * 4 + 3] - (Wm[r * 4 + 1] * Cp[3] + Wm[r * 4 + 3] * Cp[8]);
        }
        if (has_pred) inv4x4(Dp, Di);
        if (it == 0) {
            // a chain head's T row reaches this node's successor: X = -W(:,T) * Eout_head extends Cout
            const double e0 = shfl_from(Eout[0], ch.pred), e1 = shfl_from(Eout[1], ch.pred), e2 = shfl_from(Eout[2], ch.pred);
            const bool p1 = ch.pos == 1;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                Xo[r * 3 + 0] = p1 ? -Wm[r * 4 + 3] * e0 : 0.0;
                Xo[r * 3 + 1] = p1 ? -Wm[r * 4 + 3] * e1 : 0.0;
                Xo[r * 3 + 2] = p1 ? -Wm[r * 4 + 3] * e2 : 0.0;
            }
#pragma unroll
            for (int k = 0; k < 12; k++) Xp[k] = shfl_from(Xo[k], ch.pred);
        }
    }
    // P = Dinv * (Cout + Xo)
    double Pm[16];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const double d0 = Di[r * 4 + 0], d1 = Di[r * 4 + 1], d2 = Di[r * 4 + 2], d3 = Di[r * 4 + 3];
        Pm[r * 4 + 0] = d0 * (Cout[0] + Xo[0]) + d1 * (Cout[1] + Xo[3]) + d2 * Xo[6] + d3 * (Cout[5] + Xo[9]);
        Pm[r * 4 + 1] = d0 * Xo[1] + d1 * (Cout[2] + Xo[4]) + d2 * Xo[7] + d3 * (Cout[6] + Xo[10]);
        Pm[r * 4 + 2] = d0 * Xo[2] + d1 * Xo[5] + d2 * (Cout[4] + Xo[8]) + d3 * (Cout[7] + Xo[11]);
        Pm[r * 4 + 3] = d1 * Cout[3] + d3 * Cout[8];
    }
    {   // meeting node: second neighbour (mid+1, right chain): Wr = U_mid Dinv_{mid+1}, stored in Pm
        const int mid = m.mid;
        double G[16], q9[9];
#pragma unroll
        for (int k = 0; k < 16; k++) G[k] = shfl_from(Di[k], mid + 1);
#pragma unroll
        for (int k = 0; k < 9; k++) q9[k] = shfl_from(L9[k], mid + 1);
        if (ch.is_mid) {
            double Wr[16];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                Wr[c] = U9[0] * G[c];
                Wr[4 + c] = U9[1] * G[c] + U9[2] * G[4 + c] + U9[3] * G[12 + c];
                Wr[8 + c] = U9[4] * G[8 + c];
                Wr[12 + c] = U9[5] * G[c] + U9[6] * G[4 + c] + U9[7] * G[8 + c] + U9[8] * G[12 + c];
            }
            double Dp[16];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                Dp[r * 4 + 0] = Dm[r * 4 + 0] - (Wm[r * 4 + 0] * Cp[0] + Wm[r * 4 + 1] * Cp[1] + Wm[r * 4 + 3] * Cp[5])
                                - (Wr[r * 4 + 0] * q9[0] + Wr[r * 4 + 1] * q9[1] + Wr[r * 4 + 3] * q9[5]);
                Dp[r * 4 + 1] = Dm[r * 4 + 1] - (Wm[r * 4 + 1] * Cp[2] + Wm[r * 4 + 3] * Cp[6]) - (Wr[r * 4 + 1] * q9[2] + Wr[r * 4 + 3] * q9[6]);
                Dp[r * 4 + 2] = Dm[r * 4 + 2] - (Wm[r * 4 + 2] * Cp[4] + Wm[r * 4 + 3] * Cp[7]) - (Wr[r * 4 + 2] * q9[4] + Wr[r * 4 + 3] * q9[7]);
                Dp[r * 4 + 3] = Dm[r * 4 + 3] - (Wm[r * 4 + 1] * Cp[3] + Wm[r * 4 + 3] * Cp[8]) - (Wr[r * 4 + 1] * q9[3] + Wr[r * 4 + 3] * q9[8]);
            }
            inv4x4(Dp, Di);
#pragma unroll
            for (int k = 0; k < 16; k++) Pm[k] = Wr[k];
        }
    }
#if PLB_TH_BLOCKS_GLOBAL
    {
        double* const gb = Fa.blk;
#pragma unroll
        for (int k = 0; k < 16; k++) { gb[k * LW + lane] = Di[k]; gb[(16 + k) * LW + lane] = Wm[k]; gb[(32 + k) * LW + lane] = Pm[k]; }
    }
#else
#pragma unroll
    for (int k = 0; k < 16; k++) { Fa.Dinv[k][lane] = Di[k]; Fa.Wm[k][lane] = Wm[k]; Fa.Pm[k][lane] = Pm[k]; }
#endif
#pragma unroll
    for (int k = 0; k < 4; k++) Fa.Fr[k][lane] = Fr[k];
#pragma unroll
    for (int k = 0; k < 3; k++) Fa.Eo[k][lane] = Eout[k];
    grp_sync();
    // ---- 5. border: z = core^{-1} (dF/dI column), then the Schur complement ------------------------
#if PLB_SEI
    // the side-reaction rows depend on I (rate law ~ I^w): their column entry, eliminated node-locally
    const double zb[4] = {Sj[0] * QIv[0] + Sj[1] * QIv[1] + Sj[2] * QIv[2], Sj[3] * QIv[0] + Sj[4] * QIv[1] + Sj[5] * QIv[2],
                          J.ps_I + Sj[6] * QIv[0] + Sj[7] * QIv[1] + Sj[8] * QIv[2], Sj[9] * QIv[0] + Sj[10] * QIv[1] + Sj[11] * QIv[2]};
#else
    const double zb[4] = {0.0, 0.0, J.ps_I, 0.0};
#endif
    double u4[4], ux;
    core_solve(m, ch, Fa, Di, Wm, Pm, zb, (ch.ch && dyn) ? J.Tx_I : 0.0, u4, ux, lane);
#pragma unroll
    for (int k = 0; k < 4; k++) Fa.z[k][lane] = u4[k];
    Fa.zx[lane] = ux;
    const bool isdt = ctrl.gTn != 0.0 || ctrl.gTx != 0.0;     // same on every active lane
    const int mode = (grp_max(isdt ? 1.0 : 0.0) != 0.0) ? (dyn ? 1 : 2) : 0;
    Fa.gT[lane] = cj * ctrl.gTn; Fa.gX[lane] = cj * ctrl.gTx;
    if (mode == 2) {
        // d(control row)/d(algebraic unknowns) = sum_x gTn[x] * (T row of node x); the collector rows only see I
        FA_PD(0) = ctrl.gTn * J.T_j;
#pragma unroll
        for (int k = 0; k < 5; k++) FA_PD(1 + k) = ctrl.gTn * J.T_pe[k];
#pragma unroll
        for (int k = 0; k < 4; k++) FA_PD(6 + k) = ctrl.gTn * J.T_ps[k];
        FA_WT(0) = ctrl.gTn * J.T_ps[4];
#if PLB_SEI
        Fa.pdjs[lane] = ctrl.gTn * J.T_js;
#endif
    }
    if (lane == 0) { Fa.g_ps0 = ctrl.g_ps0; Fa.g_psN = ctrl.g_psN; Fa.mode = (double)mode; Fa.g_eta = ctrl.g_eta; }
    grp_sync();
#if PLB_SEI
    // column solution of the local unknowns: l_z = Mi (cI - Cl u_z) = Ql u_z - QI
    const double djz = Ql[0] * u4[0] + Ql[1] * u4[1] + Ql[2] * u4[2] + Ql[3] * u4[3] - QIv[0];
    const double djsz = Ql[4] * u4[0] + Ql[5] * u4[1] + Ql[6] * u4[2] + Ql[7] * u4[3] - QIv[1];
#else
    // the border column has no entry in the j rows: dj of the column solution is q . z
    const double djz = ro.elec ? q[0] * u4[0] + q[1] * u4[1] + q[2] * u4[2] + q[3] * u4[3] : 0.0;
    const double djsz = 0.0;
#endif
    const double gz = border_dot(m, Fa, mode, u4, ux, djz, djsz, lane);
    if (lane == 0) Fa.schur_inv = 1.0 / (ctrl.g_I - gz);
    grp_sync();
}

// Solve J * d = g for one right-hand side held node-wise in registers (g in, d out, in place).
__device__ __forceinline__ double warp_solve_impl(const ModelDesc& m, const LaneRole& ro, const WarpFactor& Fa,
                                                  bool alg_only, LaneVec& g, double gI, int lane) {
    const LaneChain ch = make_chain(m, ro, lane);
    const bool dyn = !alg_only;
    // particle right-hand side into the eigen-basis: w0 = EVI * g_cs ; s9 = surface component of A^{-1} g_cs
    double s9 = 0.0;
    if (dyn) {
        double w0[NR];
#pragma unroll
        for (int i = 0; i < NR; i++) {
            double w = 0.0;
#pragma unroll
            for (int c = 0; c < NR; c++) w = fma(laws::EVI[i][c], g.cs[c], w);
            w0[i] = ro.elec ? w : 0.0;
            s9 = fma(laws::EV[NR - 1][i] * FA_PD(i), w0[i], s9);
        }
#pragma unroll
        for (int i = 0; i < NR; i++) g.cs[i] = w0[i];
    }
    double rb[4];
#if PLB_SEI
    const bool sn = ro.sec == 2;
    const double v0 = ro.elec ? g.j - Fa.jcs[lane] * s9 : 0.0, v1 = sn ? g.js : 0.0, v2 = (sn && dyn) ? g.film : 0.0;
    double l0[3];     // Mi v
#pragma unroll
    for (int r = 0; r < 3; r++) l0[r] = Fa.Mi[r * 3 + 0][lane] * v0 + Fa.Mi[r * 3 + 1][lane] * v1 + Fa.Mi[r * 3 + 2][lane] * v2;
    rb[0] = dyn ? g.ce - (Fa.SM[0][lane] * v0 + Fa.SM[1][lane] * v1 + Fa.SM[2][lane] * v2) : 0.0;
    rb[1] = g.pe - (Fa.SM[3][lane] * v0 + Fa.SM[4][lane] * v1 + Fa.SM[5][lane] * v2);
    rb[2] = (ro.elec ? g.ps : 0.0) - (Fa.SM[6][lane] * v0 + Fa.SM[7][lane] * v1 + Fa.SM[8][lane] * v2);
    rb[3] = dyn ? g.T - (Fa.SM[9][lane] * v0 + Fa.SM[10][lane] * v1 + Fa.SM[11][lane] * v2) - Fa.tcs[lane] * s9 : 0.0;
#else
    const double q0 = ro.elec ? (g.j - Fa.jcs[lane] * s9) * Fa.q[4][lane] : 0.0;
    rb[0] = dyn ? g.ce - Fa.sj[0][lane] * q0 : 0.0;
    rb[1] = g.pe - Fa.sj[1][lane] * q0;
    rb[2] = (ro.elec ? g.ps : 0.0) - Fa.sj[2][lane] * q0;
    rb[3] = dyn ? g.T - Fa.sj[3][lane] * q0 - Fa.tcs[lane] * s9 : 0.0;
#endif
    if (!ro.act) { rb[0] = rb[1] = rb[2] = rb[3] = 0.0; }
    double Di[16], Wm[16], Pm[16];
#if PLB_TH_BLOCKS_GLOBAL
    {
        const double* const gb = Fa.blk;
#pragma unroll
        for (int k = 0; k < 16; k++) { Di[k] = gb[k * LW + lane]; Wm[k] = gb[(16 + k) * LW + lane]; Pm[k] = gb[(32 + k) * LW + lane]; }
    }
#else
#pragma unroll
    for (int k = 0; k < 16; k++) { Di[k] = Fa.Dinv[k][lane]; Wm[k] = Fa.Wm[k][lane]; Pm[k] = Fa.Pm[k][lane]; }
#endif
    double u4[4], ux;
    core_solve(m, ch, Fa, Di, Wm, Pm, rb, (ch.ch && dyn) ? g.Tx : 0.0, u4, ux, lane);
    // border
    const int mode = (int)Fa.mode;
#if PLB_SEI
    const double dj0 = l0[0] + Fa.Ql[0][lane] * u4[0] + Fa.Ql[1][lane] * u4[1] + Fa.Ql[2][lane] * u4[2] + Fa.Ql[3][lane] * u4[3];
    const double djs0 = l0[1] + Fa.Ql[4][lane] * u4[0] + Fa.Ql[5][lane] * u4[1] + Fa.Ql[6][lane] * u4[2] + Fa.Ql[7][lane] * u4[3];
#else
    const double dj0 = ro.elec ? q0 + Fa.q[0][lane] * u4[0] + Fa.q[1][lane] * u4[1] + Fa.q[2][lane] * u4[2] + Fa.q[3][lane] * u4[3] : 0.0;
    const double djs0 = 0.0;
#endif
    const double dI = (gI - border_dot(m, Fa, mode, u4, ux, dj0, djs0, lane)) * Fa.schur_inv;
#pragma unroll
    for (int k = 0; k < 4; k++) u4[k] -= Fa.z[k][lane] * dI;
    ux -= Fa.zx[lane] * dI;
    // back-substitute j (j_s, film) and the particle
#if PLB_SEI
    double lf[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
        lf[r] = l0[r] + Fa.Ql[r * 4 + 0][lane] * u4[0] + Fa.Ql[r * 4 + 1][lane] * u4[1] + Fa.Ql[r * 4 + 2][lane] * u4[2] +
                Fa.Ql[r * 4 + 3][lane] * u4[3] + Fa.QI[r][lane] * dI;
    const double dj = ro.elec ? lf[0] : 0.0;
    const double djs = sn ? lf[1] : 0.0;
    // SOH: its column only holds -cj on the diagonal; its row is dense over j_s (residuals.jl:278-297)
    const double sj_sum = warp_sum(Fa.sohc[lane] * djs);
    g.soh = dyn ? (g.soh - sj_sum) / (-Fa.cjv) : 0.0;
    g.js = djs;
    g.film = (sn && dyn) ? lf[2] : 0.0;
#else
    const double dj = ro.elec ? q0 + Fa.q[0][lane] * u4[0] + Fa.q[1][lane] * u4[1] + Fa.q[2][lane] * u4[2] + Fa.q[3][lane] * u4[3] : 0.0;
#endif
    if (dyn) {
        double v[NR];
        const double bj = Fa.csj[lane] * dj;
#pragma unroll
        for (int i = 0; i < NR; i++)
            v[i] = FA_PD(i) * (g.cs[i] - PLB_EVIB(i) * bj - FA_WT(i) * u4[3]);
#pragma unroll
        for (int r = 0; r < NR; r++) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < NR; i++) acc = fma(laws::EV[r][i], v[i], acc);
            g.cs[r] = ro.elec ? acc : 0.0;
        }
    } else {
#pragma unroll
        for (int r = 0; r < NR; r++) g.cs[r] = 0.0;
    }
    g.ce = dyn ? u4[0] : 0.0;
    g.pe = u4[1];
    g.ps = ro.elec ? u4[2] : 0.0;
    g.T = dyn ? u4[3] : 0.0;
    g.Tx = dyn ? ux : 0.0;
    g.j = dj;
    return dI;
}
#endif  // PLB_TH

// out-of-line copies for the operator-level Newton-init kernel
__device__ __noinline__ void warp_factor(const ModelDesc& m, const LaneRole& ro, const LaneJac& J,
                                         const CtrlRow& ctrl, double cj, bool alg_only, WarpFactor& Fa, int lane) {
    warp_factor_impl(m, ro, J, ctrl, cj, alg_only, Fa, lane);
}
__device__ __noinline__ double warp_solve(const ModelDesc& m, const LaneRole& ro, const WarpFactor& Fa,
                                          bool alg_only, LaneVec& g, double gI, int lane) {
    return warp_solve_impl(m, ro, Fa, alg_only, g, gI, lane);
}

}  // namespace PLB_NS
}  // namespace plb
