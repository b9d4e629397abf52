// plb_integrator.cuh -- fused per-warp DAE integrator: algebraic initialisation, variable-order
// variable-step BDF (fixed-leading-coefficient form, orders 1..5) with modified Newton, stop
// conditions with linear back-interpolation, output capture.
//
// Replaces, for one system per warp and without leaving the SM:
//   newtons_method!                 /root/reference/src/model_evaluation.jl:430-480
//   IDA (Sundials.jl `step!`)       call sites model_evaluation.jl:249-250, 262-280, 320
//                                   (third-party SUNDIALS 5.x IDA algorithm; decision logic restated
//                                   so that step/order sequences coincide with the reference's)
//   solve! / check_simulation_stop! / check_solve / interp_final_points!
//                                   model_evaluation.jl:312-382, checks.jl:1-249
//   set_vars! (t, V, I, SOC rows)   save_outputs.jl:11-40
#pragma once
#include <float.h>

#include "plb_device.cuh"

namespace plb {
namespace PLB_NS {

// vector stride in doubles (N_tot padded): 301 / 351 / 322 on the 32-node families, up to 642 (N=(20,20,20) with SEI) wide
// (N_r = 12 / 14 sibling builds: (N_r - 10) more doubles per electrode node -- 20 / 40 nodes of the standard grids; 8-aligned)
constexpr int VS = (WIDE ? (TH ? (SEI ? 784 : 736) : 656) : (TH ? (SEI ? 384 : 352) : (SEI ? 336 : 304))) + (NR - 10) * (WIDE ? 40 : 20);
// (Round 2: no predictor vectors.  y_pred = sum_j phi_j and y'_pred = sum_j gamma_j phi_j are re-formed from the
// history where an evaluation needs them -- kk+1 shared-memory reads per component instead of two: eight vectors per system
// instead of ten, of which the last PLB_NGLOBAL history vectors may live in global memory, below.)
enum VecId { V_PHI0 = 0, V_PHI1, V_PHI2, V_PHI3, V_PHI4, V_PHI5, V_EWT, V_EE, V_COUNT };

struct IdaCoef {
    double psi[6], alpha[6], beta[6], sigma[6], gamma[6];
    double cvals[6], dvals[6];     // interpolation weights at the return time
    double cprev[6];               // interpolation weights at the previous return time
};

// Number of BDF history vectors kept in global memory (L2-resident) instead of shared memory: the
// LAST PLB_NGLOBAL of phi_0..phi_5.  Measured order histogram of the 1C discharge batch: order 1: 6 %,
// 2: 47 %, 3: 41 %, 4: 5 %, 5: 0.1 % -- phi_5 is hardly ever touched, phi_4 is written once per step at order 3
// (a fire-and-forget store) and read when an order raise is considered.  Parking phi_4 and phi_5 there costs 10 % at equal
// occupancy (first sitting of round 2: it bought 8 instead of 6 systems per SM).  Second sitting: the factored blocks of the
// linear solve are what belongs in the global workspace (plb_device.cuh: touched once per solve, not once per vector pass); with
// them gone the isothermal families keep ALL history vectors on chip again (iso 369 k -> 432 k sims/s at the same eight systems
// per SM), the others park what buys them one more system per SM -- the table in plb_variant.cuh (PLB_SIM_WARPS).
#ifndef PLB_NGLOBAL
#if PLB_WIDE
#define PLB_NGLOBAL (PLB_TH ? (PLB_SEI ? 0 : 1) : 1)
#else
#define PLB_NGLOBAL (PLB_TH ? (PLB_SEI ? 1 : 2) : 0)
#endif
#endif
constexpr int NGLOBAL = PLB_NGLOBAL;
constexpr int NSHARED = V_COUNT - NGLOBAL;
constexpr int GL_FIRST = 6 - NGLOBAL;   // ids [GL_FIRST, 6) are global

struct alignas(16) WarpSmem {
    double svec[NSHARED > 0 ? NSHARED : 1][VS];
    WarpConst C;
    WarpFactor Fa;
    IdaCoef K;
};

struct WarpWS {
    double* gbase;     // this system's slot of the global workspace: NGLOBAL vectors of VS doubles (then the factored blocks, Fa.blk)
    double* sbase;     // shared-memory vectors
    WarpConst& C;
    WarpFactor& Fa;
    IdaCoef& K;
    __device__ __forceinline__ double* v(int id) const {
        if (NGLOBAL > 0 && id >= GL_FIRST && id < 6) return gbase + (id - GL_FIRST) * VS;
        return sbase + (id < 6 ? id : id - NGLOBAL) * VS;
    }
};

// internal vector layout: c_e[Nx] | c_s | [T: a|p|s|n|z] | j[Ne] | Phi_e[Nx] | Phi_s[Ne] | I
// PLB_CS_PMAJOR = 1: c_s particle-major [Ne][NR] -- the reference's own order, so the state I/O is the identity map,
//   and a lane moves its ten radial values with five 128-bit accesses (conflict-free per quarter-warp: lane stride
//   80 bytes); falls back to 64-bit accesses when N_x is odd (the block then starts on an odd word).
// PLB_CS_PMAJOR = 0: radial-major [NR][Ne] (64-bit accesses, conflict-free)
#ifndef PLB_CS_PMAJOR
#define PLB_CS_PMAJOR 1
#endif
static_assert(NR % 2 == 0, "paired particle accesses");
__device__ __forceinline__ void load_lane(const ModelDesc& m, const LaneRole& ro, const double* v,
                                          LaneVec& y, double& I) {
    y.ce = ro.act ? v[ro.x] : 0.0;
    y.pe = ro.act ? v[m.off_pe + ro.x] : 0.0;
    if (ro.elec) {
#if PLB_CS_PMAJOR
        const double* pc = v + m.off_cs + ro.e * NR;
        if ((m.off_cs & 1) == 0) {
#pragma unroll
            for (int k = 0; k < NR / 2; k++) {
                const double2 t = reinterpret_cast<const double2*>(pc)[k];
                y.cs[2 * k] = t.x; y.cs[2 * k + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int r = 0; r < NR; r++) y.cs[r] = pc[r];
        }
#else
#pragma unroll
        for (int r = 0; r < NR; r++) y.cs[r] = v[m.off_cs + r * m.Ne + ro.e];
#endif
        y.j = v[m.off_j + ro.e];
        y.ps = v[m.off_ps + ro.e];
    } else {
#pragma unroll
        for (int r = 0; r < NR; r++) y.cs[r] = 0.0;
        y.j = 0.0; y.ps = 0.0;
    }
    if (TH) {
        y.T = ro.act ? v[m.off_T + m.Na + ro.x] : 0.0;
        y.Tx = ro.ix >= 0 ? v[m.off_T + ro.ix] : 0.0;
    }
    if (SEI) {
        const int k = ro.x - (m.Np + m.Ns);
        y.js = ro.sec == 2 ? v[m.off_js + k] : 0.0;
        y.film = ro.sec == 2 ? v[m.off_film + k] : 0.0;
        y.soh = v[m.off_SOH];
    }
    I = v[m.off_I];
}
__device__ __forceinline__ void store_lane(const ModelDesc& m, const LaneRole& ro, double* v,
                                           const LaneVec& y, double I, int lane) {
    if (ro.act) { v[ro.x] = y.ce; v[m.off_pe + ro.x] = y.pe; }
    if (ro.elec) {
#if PLB_CS_PMAJOR
        double* pc = v + m.off_cs + ro.e * NR;
        if ((m.off_cs & 1) == 0) {
#pragma unroll
            for (int k = 0; k < NR / 2; k++) reinterpret_cast<double2*>(pc)[k] = make_double2(y.cs[2 * k], y.cs[2 * k + 1]);
        } else {
#pragma unroll
            for (int r = 0; r < NR; r++) pc[r] = y.cs[r];
        }
#else
#pragma unroll
        for (int r = 0; r < NR; r++) v[m.off_cs + r * m.Ne + ro.e] = y.cs[r];
#endif
        v[m.off_j + ro.e] = y.j;
        v[m.off_ps + ro.e] = y.ps;
    }
    if (TH) {
        if (ro.act) v[m.off_T + m.Na + ro.x] = y.T;
        if (ro.ix >= 0) v[m.off_T + ro.ix] = y.Tx;
    }
    if (SEI) {
        const int k = ro.x - (m.Np + m.Ns);
        if (ro.sec == 2) { v[m.off_js + k] = y.js; v[m.off_film + k] = y.film; }
        if (lane == 0) v[m.off_SOH] = y.soh;
    }
    if (lane == 0) v[m.off_I] = I;
}
// reference (particle-major) layout <-> internal index
__device__ __forceinline__ int ref_index(const ModelDesc& m, int i) {
#if PLB_CS_PMAJOR
    return i;
#endif
    if (i < m.off_cs || i >= m.off_cs + NR * m.Ne) return i;
    const int k = i - m.off_cs, r = k / m.Ne, e = k % m.Ne;
    return m.off_cs + e * NR + r;
}

__device__ __forceinline__ double wrms(const ModelDesc& m, const double* v, const double* w, int lane) {
    double s = 0.0;
    for (int i = lane; i < m.N_tot; i += LW) { const double p = v[i] * w[i]; s = fma(p, p, s); }
    return sqrt(warp_sum(s) / m.N_tot);
}

struct RunCtl {
    int method;
    int dc;            // METHOD_DC: target lane | component << 8 (0: surface c_s, 1: c_e); fills the padding
    double value;
};
// the method word a lane_eval call gets (see plb_common.cuh); alg: the newtons_method! form of a Y'-dependent row
__device__ __forceinline__ int method_word(const RunCtl& rc, bool alg) {
    if (alg && rc.method == METHOD_DT) return METHOD_DT_ALG;
#if PLB_DC
    if (rc.method == METHOD_DC) return (alg ? METHOD_DC_ALG : METHOD_DC) | (rc.dc << 8);
#endif
    return rc.method;
}

// ------------------------------------------------------------------------------------------------
// newtons_method! -- model_evaluation.jl:430-480.  Y: vector in workspace (in/out), YP: vector (out)
// returns iterations (>0) or FAIL_NEWTON_INIT
// ------------------------------------------------------------------------------------------------
template <int CHEM>
__device__ __forceinline__ int newton_init(const ModelDesc& m, WarpWS& w, const LaneRole& ro,
                                           const RunCtl& rc, const Opts& o, double* Y, double* YP,
                                           int lane, int& n_res, int& n_jac) {
    LaneVec y, yp, res;
    LaneJac J;
    CtrlRow ctrl;
    double I;
    load_lane(m, ro, Y, y, I);
    yp.ce = 0.0; yp.j = 0.0; yp.pe = 0.0; yp.ps = 0.0; yp.T = 0.0; yp.Tx = 0.0; yp.js = 0.0; yp.film = 0.0; yp.soh = 0.0;
#pragma unroll
    for (int r = 0; r < NR; r++) yp.cs[r] = 0.0;
    int iter;
    bool ok = false;
    const int meth = method_word(rc, true);
    for (iter = 1; iter <= 100; iter++) {
        lane_eval_ni<CHEM, true>(m, w.C, ro, y, yp, I, meth, rc.value, res, ctrl, J);   // R_alg, J_alg
        n_res++; n_jac++;
        warp_factor(m, ro, J, ctrl, 0.0, true, w.Fa, lane);
        const double dI = warp_solve(m, ro, w.Fa, true, res, ctrl.res, lane);
        // Y_new .-= factor \ res ; stop on the absolute 2-norm of the update (:451-454)
        double s = 0.0;
        if (ro.elec) { y.j -= res.j; y.ps -= res.ps; s += res.j * res.j + res.ps * res.ps; }
        if (ro.act) { y.pe -= res.pe; s += res.pe * res.pe; }
        if (SEI && ro.sec == 2) { y.js -= res.js; s += res.js * res.js; }
        I -= dI;
        s = warp_sum(s) + dI * dI;
        if (!(s == s) || isinf(s)) return FAIL_NEWTON_INIT;
        if (sqrt(s) < o.reltol_init) { ok = true; break; }
    }
    if (!ok) return FAIL_NEWTON_INIT;
    // R_diff(YP,t,Y,YP): YP_diff = rhs of the differential rows (:460)
    lane_eval_ni<CHEM, false>(m, w.C, ro, y, yp, I, meth, rc.value, res, ctrl, J);
    n_res++;
    LaneVec ypo;
    ypo.ce = res.ce;
    ypo.T = TH ? res.T : 0.0; ypo.Tx = TH ? res.Tx : 0.0;
    ypo.film = SEI ? res.film : 0.0; ypo.soh = SEI ? res.soh : 0.0; ypo.js = 0.0;
#pragma unroll
    for (int r = 0; r < NR; r++) ypo.cs[r] = res.cs[r];
    ypo.j = 0.0; ypo.pe = 0.0; ypo.ps = 0.0;
    if (o.skip_alg_deriv) {   // initialize_algebraic_derivatives = false (:433, :462): Y'_alg stays 0
        store_lane(m, ro, Y, y, I, lane);
        store_lane(m, ro, YP, ypo, 0.0, lane);
    } else {
        // estimate dY_alg/dt (:462-477): Delta_t = max(10 reltol_init, sqrt(eps(c_e0)))
        const double c0 = fabs(w.C.theta[TF_c_e0]);
        const double epsv = ::nextafter(c0, DBL_MAX) - c0;
        const double dt = fmax(10.0 * o.reltol_init, sqrt(epsv));
        LaneVec yn = y;
        yn.ce = y.ce + dt * ypo.ce;
        if (TH) { yn.T = y.T + dt * ypo.T; yn.Tx = y.Tx + dt * ypo.Tx; }
        if (SEI) { yn.film = y.film + dt * ypo.film; yn.soh = y.soh + dt * ypo.soh; }
#pragma unroll
        for (int r = 0; r < NR; r++) yn.cs[r] = y.cs[r] + dt * ypo.cs[r];
        lane_eval_ni<CHEM, false>(m, w.C, ro, yn, yp, I, meth, rc.value, res, ctrl, J);
        n_res++;
        const double dI = warp_solve(m, ro, w.Fa, true, res, ctrl.res, lane);
        ypo.j = -res.j / dt; ypo.pe = -res.pe / dt; ypo.ps = -res.ps / dt;
        if (SEI) ypo.js = -res.js / dt;
        store_lane(m, ro, Y, y, I, lane);
        store_lane(m, ro, YP, ypo, -dI / dt, lane);
    }
    grp_sync();
    return iter;
}

// ------------------------------------------------------------------------------------------------
// IDA state (uniform across the warp, in registers) + helpers
// ------------------------------------------------------------------------------------------------
struct Ida {
    double tn, hh, hused, cj, cjlast, cjold, cjratio, ss, rr, hin;
    double tstop, tretlast;
    int kk, kused, knew, phase, ns, nst, tstopset;
    int nre, nje, netf, ncfn;
};

// ------------------------------------------------------------------------------------------------
// vector passes over the N_tot components: element i = lane + 32*k, k < NEL, fully unrolled so the
// NEL independent chains overlap (these passes sit on the latency-critical path between two
// residual evaluations).  Summation order per lane (k ascending, then the xor-tree) is fixed.
// ------------------------------------------------------------------------------------------------
constexpr int NEL = (VS + LW - 1) / LW;
#ifndef PLB_ELEM_UNROLL
#define PLB_ELEM_UNROLL 1
#endif
#define PLB_STR_(x) #x
#define PLB_STR(x) PLB_STR_(x)
#define PLB_FOR_ELEMS(i, N) _Pragma(PLB_STR(unroll PLB_ELEM_UNROLL)) for (int i = lane; i < (N); i += LW)
// the hot passes (weights, predictor, error norms, phi update) move two components per 128-bit shared-memory
// access: pair q = elements 2q, 2q+1.  For odd N the last pair reaches into the padding of the vectors (VS >= N+1),
// which the integrator zeroes once per launch and which stays zero (its weight is forced to zero).
#ifndef PLB_VEC2
#define PLB_VEC2 1
#endif
static_assert(VS % 2 == 0 && (VS * sizeof(double)) % 16 == 0, "vector stride must keep 16-byte alignment");
#define PLB_FOR_PAIRS(q, N) _Pragma(PLB_STR(unroll PLB_ELEM_UNROLL)) for (int q = lane; q < ((N) + 1) / 2; q += LW)
__device__ __forceinline__ double2 ld2(const double* v, int q) { return reinterpret_cast<const double2*>(v)[q]; }
__device__ __forceinline__ void st2(double* v, int q, double2 x) { reinterpret_cast<double2*>(v)[q] = x; }

__device__ __forceinline__ void ewt_set(const ModelDesc& m, WarpWS& w, const Opts& o, int lane) {
    const double* p0 = w.v(V_PHI0);
    double* ew = w.v(V_EWT);
    const double rt = o.reltol, at = o.abstol;
#if PLB_VEC2
    PLB_FOR_PAIRS(q, m.N_tot) {
        const double2 p = ld2(p0, q);
        double2 e;
        e.x = 1.0 / (rt * fabs(p.x) + at);
        e.y = (2 * q + 1 < m.N_tot) ? 1.0 / (rt * fabs(p.y) + at) : 0.0;
        st2(ew, q, e);
    }
#else
    PLB_FOR_ELEMS(i, m.N_tot) ew[i] = 1.0 / (rt * fabs(p0[i]) + at);
#endif
    grp_sync();
}

// interpolation weights of IDAGetSolution at time t -> c[0..kord], d[0..kord-1]
// (delt = t - tn.  For the previous step point pass -hused itself: formed as t_prev - tn it loses h's low bits, which
// matters once h has shrunk to ~1e-8 s in front of a failure -- the weights of phi_2.. are then not exactly zero.)
__device__ __forceinline__ int getsol_weights_delt(const Ida& M, const IdaCoef& K, double delt, double* c, double* d) {
    int kord = M.kused; if (kord == 0) kord = 1;
    double cc = 1.0, dd = 0.0, gam = delt / K.psi[0];
    c[0] = cc;
#pragma unroll 1
    for (int j = 1; j <= kord; j++) {
        dd = dd * gam + cc / K.psi[j - 1];
        cc = cc * gam;
        gam = (delt + K.psi[j - 1]) / K.psi[j];
        c[j] = cc; d[j - 1] = dd;
    }
    return kord;
}
__device__ __forceinline__ int getsol_weights(const Ida& M, const IdaCoef& K, double t, double* c, double* d) {
    return getsol_weights_delt(M, K, t - M.tn, c, d);
}

// IDASetCoeffs, scalar part (the phi scaling is fused into predict_pass).  Returns ck.
__device__ __forceinline__ double ida_set_coeffs(const ModelDesc& m, WarpWS& w, Ida& M, int lane) {
    IdaCoef& K = w.K;
    if (M.hh != M.hused || M.kk != M.kused) M.ns = 0;
    M.ns = min(M.ns + 1, M.kused + 2);
    grp_sync();
    if (M.kk + 1 >= M.ns) {
        if (lane == 0) {
            K.beta[0] = 1.0; K.alpha[0] = 1.0; K.gamma[0] = 0.0; K.sigma[0] = 1.0;
            double temp1 = M.hh;
#pragma unroll 1
            for (int i = 1; i <= M.kk; i++) {
                const double temp2 = K.psi[i - 1];
                K.psi[i - 1] = temp1;
                K.beta[i] = K.beta[i - 1] * K.psi[i - 1] / temp2;
                temp1 = temp2 + M.hh;
                K.alpha[i] = M.hh / temp1;
                K.sigma[i] = i * K.sigma[i - 1] * K.alpha[i];
                K.gamma[i] = K.gamma[i - 1] + K.alpha[i - 1] / M.hh;
            }
            K.psi[M.kk] = temp1;
        }
    }
    grp_sync();
    double alphas = 0.0, alpha0 = 0.0;
#pragma unroll 1
    for (int i = 0; i < M.kk; i++) { alphas -= 1.0 / (i + 1); alpha0 -= K.alpha[i]; }
    M.cjlast = M.cj;
    M.cj = -alphas / M.hh;
    double ck = fabs(K.alpha[M.kk] + alphas - alpha0);
    ck = fmax(ck, K.alpha[M.kk]);
    M.tn += M.hh;
    return ck;
}

// phi_k *= beta_k (k = ns..kk: "phi-star"), ee = 0.  The predictor itself (y = sum phi_j, y' = sum gamma_j phi_j) is
// formed per lane by predictor_lane() where an evaluation needs it.
__device__ __forceinline__ void predict_pass(const ModelDesc& m, WarpWS& w, const Ida& M, int lane) {
    const IdaCoef& K = w.K;
    double be[6];
#pragma unroll
    for (int j = 0; j < 6; j++) be[j] = K.beta[j];
    double* ee_ = w.v(V_EE);
    const int kk = M.kk, ns = M.ns;     // the integrator state lives in shared memory: read it once
#if PLB_VEC2
    PLB_FOR_PAIRS(q, m.N_tot) {
#pragma unroll
        for (int j = 1; j < 6; j++) {
            if (j >= ns && j <= kk) {
                double* ph = w.v(V_PHI0 + j);
                double2 p = ld2(ph, q);
                p.x *= be[j]; p.y *= be[j];
                st2(ph, q, p);
            }
        }
        st2(ee_, q, make_double2(0.0, 0.0));
    }
#else
    PLB_FOR_ELEMS(i, m.N_tot) {
#pragma unroll
        for (int j = 1; j < 6; j++) {
            if (j >= ns && j <= kk) { double* ph = w.v(V_PHI0 + j); ph[i] *= be[j]; }
        }
        ee_[i] = 0.0;
    }
#endif
    grp_sync();
}

// y_pred, y'_pred of this lane's components from the (phi-star) history, in the summation order the predictor
// vectors used to be formed in: y = ((phi_0 + phi_1) + ...), y' = fma(gamma_j, phi_j, y') for j = 1..kk
__device__ __forceinline__ void predictor_lane(const ModelDesc& m, const LaneRole& ro, WarpWS& w, int kk, LaneVec& y,
                                               double& Iy, LaneVec& yp) {
    load_lane(m, ro, w.v(V_PHI0), y, Iy);
    yp.ce = 0.0; yp.j = 0.0; yp.pe = 0.0; yp.ps = 0.0; yp.T = 0.0; yp.Tx = 0.0; yp.js = 0.0; yp.film = 0.0; yp.soh = 0.0;
#pragma unroll
    for (int r = 0; r < NR; r++) yp.cs[r] = 0.0;
#pragma unroll 1
    for (int j = 1; j <= kk; j++) {
        const double ga = w.K.gamma[j];
        LaneVec p;
        double pI;
        load_lane(m, ro, w.v(V_PHI0 + j), p, pI);
        y.ce += p.ce; y.j += p.j; y.pe += p.pe; y.ps += p.ps; Iy += pI;
        yp.ce = fma(ga, p.ce, yp.ce);
        if (TH) { y.T += p.T; y.Tx += p.Tx; yp.T = fma(ga, p.T, yp.T); yp.Tx = fma(ga, p.Tx, yp.Tx); }
        if (SEI) {
            y.js += p.js; y.film += p.film; y.soh += p.soh;
            yp.film = fma(ga, p.film, yp.film); yp.soh = fma(ga, p.soh, yp.soh);
        }
#pragma unroll
        for (int r = 0; r < NR; r++) { y.cs[r] += p.cs[r]; yp.cs[r] = fma(ga, p.cs[r], yp.cs[r]); }
    }
}

__device__ __forceinline__ bool ida_test_error(const ModelDesc& m, WarpWS& w, Ida& M, double ck,
                                               double& err_k, double& err_km1, int lane) {
    const IdaCoef& K = w.K;
    const double* ee = w.v(V_EE);
    const double* ewt = w.v(V_EWT);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const double* pk = w.v(V_PHI0 + M.kk);
    const double* pk1 = w.v(V_PHI0 + (M.kk > 0 ? M.kk - 1 : 0));
#if PLB_VEC2
    PLB_FOR_PAIRS(q, m.N_tot) {
        const double2 e = ld2(ee, q), wt = ld2(ewt, q), k0 = ld2(pk, q), k1 = ld2(pk1, q);
        const double ax = e.x * wt.x, ay = e.y * wt.y;
        s0 = fma(ax, ax, s0); s0 = fma(ay, ay, s0);
        const double d1x = k0.x + e.x, d1y = k0.y + e.y;
        const double bx = d1x * wt.x, by = d1y * wt.y;
        s1 = fma(bx, bx, s1); s1 = fma(by, by, s1);
        const double cx = (d1x + k1.x) * wt.x, cy = (d1y + k1.y) * wt.y;
        s2 = fma(cx, cx, s2); s2 = fma(cy, cy, s2);
    }
#else
    PLB_FOR_ELEMS(i, m.N_tot) {
        const double e = ee[i], wt = ewt[i];
        const double a = e * wt; s0 = fma(a, a, s0);
        const double d1 = (pk[i] + e); const double b = d1 * wt; s1 = fma(b, b, s1);
        const double d2 = (d1 + pk1[i]); const double c = d2 * wt; s2 = fma(c, c, s2);
    }
#endif
    const double enorm_k = sqrt(warp_sum(s0) / m.N_tot);
    err_k = K.sigma[M.kk] * enorm_k;
    const double terr_k = (M.kk + 1) * err_k;
    M.knew = M.kk;
    if (M.kk > 1) {
        err_km1 = K.sigma[M.kk - 1] * sqrt(warp_sum(s1) / m.N_tot);
        const double terr_km1 = M.kk * err_km1;
        if (M.kk > 2) {
            const double err_km2 = K.sigma[M.kk - 2] * sqrt(warp_sum(s2) / m.N_tot);
            const double terr_km2 = (M.kk - 1) * err_km2;
            if (fmax(terr_km1, terr_km2) <= terr_k) M.knew = M.kk - 1;
        } else {
            if (terr_km1 <= 0.5 * terr_k) M.knew = M.kk - 1;
        }
    }
    return ck * enorm_k > 1.0;
}

__device__ __forceinline__ void ida_restore(const ModelDesc& m, WarpWS& w, Ida& M, double saved_t, int lane) {
    IdaCoef& K = w.K;
    M.tn = saved_t;
    grp_sync();
    if (lane == 0) {
#pragma unroll 1
        for (int j = 1; j <= M.kk; j++) K.psi[j - 1] = K.psi[j] - M.hh;
    }
#pragma unroll
    for (int j = 0; j < 6; j++) {
        if (j >= M.ns && j <= M.kk) {
            const double sc = 1.0 / K.beta[j];
            double* ph = w.v(V_PHI0 + j);
            PLB_FOR_ELEMS(i, m.N_tot) ph[i] *= sc;
        }
    }
    grp_sync();
}

__device__ __forceinline__ void ida_complete_step(const ModelDesc& m, WarpWS& w, const Opts& o, Ida& M,
                                                  double err_k, double err_km1, int lane) {
    const double* ee = w.v(V_EE);
    const double* ewt = w.v(V_EWT);
    __syncwarp();   // M (part of the shared-memory state) is read-modify-written by all lanes together
    M.nst++;
    const int kdiff = M.kk - M.kused;
    M.kused = M.kk;
    M.hused = M.hh;
    if (M.knew == M.kk - 1 || M.kk == o.maxord) M.phase = 1;
    if (M.phase == 0) {
        if (M.nst > 1) { M.kk++; M.hh = 2.0 * M.hh; }
    } else {
        int action = 0;   // 0 unset, 1 lower, 2 maintain, 3 raise
        double err_kp1 = 0.0, err_knew;
        if (M.knew == M.kk - 1) action = 1;
        else if (M.kk == o.maxord) action = 2;
        else if (M.kk + 1 >= M.ns || kdiff == 1) action = 2;
        if (action == 0) {
            double s = 0.0;
            const double* pk = w.v(V_PHI0 + M.kk + 1);
            PLB_FOR_ELEMS(i, m.N_tot) { const double a = (ee[i] - pk[i]) * ewt[i]; s = fma(a, a, s); }
            const double enorm = sqrt(warp_sum(s) / m.N_tot);
            err_kp1 = enorm / (M.kk + 2);
            const double terr_k = (M.kk + 1) * err_k, terr_kp1 = (M.kk + 2) * err_kp1;
            if (M.kk == 1) {
                action = (terr_kp1 >= 0.5 * terr_k) ? 2 : 3;
            } else {
                const double terr_km1 = M.kk * err_km1;
                if (terr_km1 <= fmin(terr_k, terr_kp1)) action = 1;
                else if (terr_kp1 >= terr_k) action = 2;
                else action = 3;
            }
        }
        if (action == 3) { M.kk++; err_knew = err_kp1; }
        else if (action == 1) { M.kk--; err_knew = err_km1; }
        else err_knew = err_k;
        double hnew = M.hh;
        M.rr = pow(2.0 * err_knew + 1e-4, -1.0 / (M.kk + 1));
        if (M.rr >= 2.0) hnew = 2.0 * M.hh;
        else if (M.rr <= 1.0) { M.rr = fmax(0.5, fmin(0.9, M.rr)); hnew = M.hh * M.rr; }
        M.hh = hnew;
    }
    // phi updates: phi[kused+1] = ee ; phi[kused] += ee ; phi[j] += phi[j+1] (j = kused-1..0)
    {
        const int ku = M.kused;         // (shared-memory state: read once)
        double* const pnew = (ku < o.maxord) ? w.v(V_PHI0 + ku + 1) : nullptr;
#if PLB_VEC2
        PLB_FOR_PAIRS(q, m.N_tot) {
            const double2 e = ld2(ee, q);
            if (pnew) st2(pnew, q, e);
            double2 acc = e;
#pragma unroll 1
            for (int j = ku; j >= 0; j--) {
                double* ph = w.v(V_PHI0 + j);
                const double2 p = ld2(ph, q);
                acc.x = p.x + acc.x; acc.y = p.y + acc.y;
                st2(ph, q, acc);
            }
        }
#else
        PLB_FOR_ELEMS(i, m.N_tot) {
            const double e = ee[i];
            if (pnew) pnew[i] = e;
            double acc = e;
#pragma unroll 1
            for (int j = ku; j >= 0; j--) {
                double* ph = w.v(V_PHI0 + j);
                acc = ph[i] + acc;
                ph[i] = acc;
            }
        }
#endif
    }
    grp_sync();
}

// interpolated value / derivative of component i (internal index) with weights (c, d), order kord
__device__ __forceinline__ double interp_y(const WarpWS& w, const double* c, int kord, int i) {
    double y = 0.0;
#pragma unroll 1
    for (int j = 0; j <= kord; j++) y = fma(c[j], w.v(V_PHI0 + j)[i], y);
    return y;
}
__device__ __forceinline__ double interp_yp(const WarpWS& w, const double* d, int kord, int i) {
    double y = 0.0;
#pragma unroll 1
    for (int j = 1; j <= kord; j++) y = fma(d[j - 1], w.v(V_PHI0 + j)[i], y);
    return y;
}

struct PrevVals {   // boundary_stop_prev_values, structures.jl:174-184
    double frac, V, SOC, c_s_n, I, eta_plating, c_e_min, T, dfilm;
};

// temperature_weighting(T): length-weighted mean over the five sections
// (auxiliary_states_and_coefficients.jl:649-676) of a vector given by interpolation weights (c, kord);
// deriv: weights d of the derivative instead.  Isothermal variant: T0.
__device__ __forceinline__ double weighted_T(const ModelDesc& m, const WarpWS& w, const double* c, int kord,
                                             bool deriv, int lane) {
#if PLB_TH
    const int NT = m.Na + m.Nx + m.Nz;
    double s = 0.0;
    for (int i = lane; i < NT; i += LW) {
        const int x = i - m.Na;
        const int q = i < m.Na ? 0 : (x < m.Np ? 1 : (x < m.Np + m.Ns ? 2 : (x < m.Nx ? 3 : 4)));
        double v = 0.0;
        for (int j = deriv ? 1 : 0; j <= kord; j++) v = fma(c[deriv ? j - 1 : j], w.v(V_PHI0 + j)[m.off_T + i], v);
        s = fma(v, w.C.s5[0][q], s);
    }
    s = warp_sum(s);
    const double* th = w.C.theta;
    return s / (th[TF_l_a] + th[TF_l_p] + th[TF_l_s] + th[TF_l_n] + th[TF_l_z]);
#else
    return deriv ? 0.0 : w.C.g[GC_T];
#endif
}

// check_simulation_stop! -- checks.jl:1-224 (no SEI: the dfilm check is inactive; T only when thermal)
__device__ __noinline__ void check_stop(const ModelDesc& m, const WarpWS& w, const RunCtl& rc,
                                        const Opts& o, const Bounds& b, bool is_rest, double tf,
                                        PrevVals& pv, int& flag, double t, const double* c,
                                        const double* d, int kord, double SOC, double Ic, double V, int lane) {
    const double eps = t < 1.0 ? o.reltol : 0.0;
    if (t >= tf) { flag = 0; return; }
    if (!o.check_bounds || is_rest) return;
    const int iP0 = m.off_ps, iPN = m.off_ps + m.Ne - 1;
    if (rc.method != METHOD_I) {   // check_stop_I :31-54
        const double dIc = interp_yp(w, d, kord, m.off_I);
        if ((Ic - b.I_max > eps) && dIc > 0) {
            const double tf_ = (pv.I - b.I_max) / (pv.I - Ic);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 7; }
        } else if ((b.I_min - Ic > eps) && dIc < 0) {
            const double tf_ = (pv.I - b.I_min) / (pv.I - Ic);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 8; }
        }
        pv.I = Ic;
    }
    if (rc.method != METHOD_V) {   // check_stop_V :56-80 (the derivative is only needed past a bound)
        if (b.V_min - V > eps) {
            const double dV = interp_yp(w, d, kord, iP0) - interp_yp(w, d, kord, iPN);
            if (dV < 0) {
                const double tf_ = (pv.V - b.V_min) / (pv.V - V);
                if (tf_ < pv.frac) { pv.frac = tf_; flag = 1; }
            }
        } else if (V - b.V_max > eps) {
            const double dV = interp_yp(w, d, kord, iP0) - interp_yp(w, d, kord, iPN);
            if (dV > 0) {
                const double tf_ = (pv.V - b.V_max) / (pv.V - V);
                if (tf_ < pv.frac) { pv.frac = tf_; flag = 2; }
            }
        }
        pv.V = V;
    }
    // check_stop_SOC :82-104
    if ((b.SOC_min - SOC > eps) && Ic < 0) {
        const double tf_ = (pv.SOC - b.SOC_min) / (pv.SOC - SOC);
        if (tf_ < pv.frac) { pv.frac = tf_; flag = 3; }
    } else if ((SOC - b.SOC_max > eps) && Ic > 0) {
        const double tf_ = (pv.SOC - b.SOC_max) / (pv.SOC - SOC);
        if (tf_ < pv.frac) { pv.frac = tf_; flag = 4; }
    }
    pv.SOC = SOC;
#if PLB_TH
    // check_stop_T :106-124
    if (b.T_max == b.T_max && rc.method != METHOD_DT) {
        const double Tw = weighted_T(m, w, c, kord, false, lane);
        if (Tw - b.T_max > eps && weighted_T(m, w, d, kord, true, lane) > 0) {
            const double tf_ = (pv.T - b.T_max) / (pv.T - Tw);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 5; }
        }
        pv.T = Tw;
    }
#endif
    // check_stop_c_s_surf :141-161
    if (b.c_s_n_max == b.c_s_n_max) {
        double mx = -INFINITY;
        for (int e = m.Np + lane; e < m.Ne; e += LW) mx = fmax(mx, interp_y(w, c, kord, PLB_CS_PMAJOR ? m.off_cs + e * NR + NR - 1 : m.off_cs + (NR - 1) * m.Ne + e));
        mx = grp_max(mx);
        const double lim = b.c_s_n_max * w.C.theta[TF_c_max_n];
        if (Ic > 0 && mx - lim > eps) {
            const double tf_ = (pv.c_s_n - lim) / (pv.c_s_n - mx);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 6; }
        }
        pv.c_s_n = mx;
    }
    // check_stop_c_e :163-183
    if (b.c_e_min == b.c_e_min) {
        double mn = INFINITY;
        for (int i = lane; i < m.Nx; i += LW) mn = fmin(mn, interp_y(w, c, kord, i));
        mn = grp_min(mn);
        if (b.c_e_min - mn > eps) {
            const double tf_ = (pv.c_e_min - b.c_e_min) / (pv.c_e_min - mn);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 9; }
        }
        pv.c_e_min = mn;
    }
    // check_stop_eta_plating :185-202
    if (b.eta_plating_min == b.eta_plating_min) {
        const int in_ = m.off_ps + m.Np, ie = m.off_pe + m.Np + m.Ns;
        const double ep = interp_y(w, c, kord, in_) - interp_y(w, c, kord, ie);
        const double dep = interp_yp(w, d, kord, in_) - interp_yp(w, d, kord, ie);
        if (b.eta_plating_min - ep > eps && dep < 0) {
            const double tf_ = (pv.eta_plating - b.eta_plating_min) / (pv.eta_plating - ep);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 11; }
        }
        pv.eta_plating = ep;
    }
#if PLB_SEI
    // check_stop_dfilm :204-224 (the largest film growth rate over the anode; no derivative-sign test)
    {
        double mx = -INFINITY;
        for (int k = lane; k < m.Nn; k += LW) mx = fmax(mx, interp_yp(w, d, kord, m.off_film + k));
        mx = grp_max(mx);
        if (b.dfilm_max == b.dfilm_max && mx - b.dfilm_max > eps) {
            const double tf_ = (pv.dfilm - b.dfilm_max) / (pv.dfilm - mx);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 10; }
        }
        pv.dfilm = mx;
    }
#endif
}



}  // namespace PLB_NS
}  // namespace plb
